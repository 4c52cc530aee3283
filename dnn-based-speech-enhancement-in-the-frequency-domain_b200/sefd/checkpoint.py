"""Checkpoint compatibility with the reference (train_interface.py:101-116, 166-171):
`chkpt_N.pt` = {'model': model.state_dict(), 'optimizer': torch.optim.Adam.state_dict(), 'epoch': N}.

The model side needs nothing (the drop-in modules have the reference's state_dict keys).  The optimizer side converts
between torch.optim.Adam's per-parameter state ({'step', 'exp_avg', 'exp_avg_sq'} indexed in parameters() order) and
the flat moment buffers of sefd.train.FlatAdam / TrainStep, whose layout is the plan's parameter table."""
import torch


def adam_state_to_flat(opt_state, layout, n_flat, device="cpu"):
    """torch.optim.Adam.state_dict() -> (exp_avg [n_flat], exp_avg_sq [n_flat], step).
    layout: [(name, offset, numel, shape)] in parameters() order (Plan.params)."""
    m = torch.zeros(n_flat, dtype=torch.float32, device=device)
    v = torch.zeros(n_flat, dtype=torch.float32, device=device)
    ids = [i for grp in opt_state["param_groups"] for i in grp["params"]]
    if len(ids) != len(layout):
        raise ValueError(f"optimizer state has {len(ids)} parameters, the model {len(layout)}")
    step = 0
    for pid, (name, off, n, shape) in zip(ids, layout):
        st = opt_state["state"].get(pid)
        if st is None:
            continue                                               # parameter never stepped
        if tuple(st["exp_avg"].shape) != tuple(shape):
            raise ValueError(f"{name}: optimizer state shape {tuple(st['exp_avg'].shape)} != parameter shape {tuple(shape)}")
        m[off:off + n] = st["exp_avg"].reshape(-1).to(device=device, dtype=torch.float32)
        v[off:off + n] = st["exp_avg_sq"].reshape(-1).to(device=device, dtype=torch.float32)
        step = max(step, int(st["step"]))
    return m, v, step


def flat_to_adam_state(m, v, step, layout, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
    """Inverse: a state_dict that torch.optim.Adam(model.parameters(), lr).load_state_dict accepts."""
    state = {}
    if step > 0:
        for i, (_, off, n, shape) in enumerate(layout):
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": m[off:off + n].reshape(shape).clone().cpu(),
                        "exp_avg_sq": v[off:off + n].reshape(shape).clone().cpu()}
    group = {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": 0, "amsgrad": False, "maximize": False,
             "foreach": None, "capturable": False, "differentiable": False, "fused": None, "decoupled_weight_decay": False,
             "params": list(range(len(layout)))}
    return {"state": state, "param_groups": [group]}


def save_checkpoint(path, model, optimizer, epoch):
    """train_interface.py:166-171 with a sefd.train.FlatAdam / TrainStep (or a torch optimizer) in `optimizer`."""
    if hasattr(optimizer, "exp_avg"):
        eng = model._get_engine()
        opt = flat_to_adam_state(optimizer.exp_avg, optimizer.exp_avg_sq, optimizer.steps, eng._layout.params,
                                 optimizer.lr, optimizer.betas, optimizer.eps)
    else:
        opt = optimizer.state_dict()
    torch.save({"model": model.state_dict(), "optimizer": opt, "epoch": epoch}, path)


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """train_interface.py:108-111: returns the epoch to resume from."""
    ck = torch.load(path, map_location=map_location)
    model.load_state_dict(ck["model"])
    if optimizer is not None:
        if hasattr(optimizer, "exp_avg"):
            eng = model._get_engine()
            eng.sync()
            m, v, step = adam_state_to_flat(ck["optimizer"], eng._layout.params, eng.flat.numel(), eng.flat.device)
            optimizer.exp_avg.copy_(m)
            optimizer.exp_avg_sq.copy_(v)
            optimizer.steps = step
            if hasattr(optimizer, "_step_dev"):          # TrainStep keeps the Adam step count on the device
                optimizer._step_dev.fill_(int(step))
        else:
            optimizer.load_state_dict(ck["optimizer"])
    return ck["epoch"] + 1
