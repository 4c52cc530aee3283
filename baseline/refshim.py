"""Import shim for the UNMODIFIED reference used by bench.py's CPU arm and the boundary tests.

`baseline/_ref/` holds byte-identical copies of the reference's config.py / models.py / tools_for_model.py /
tools_for_loss.py / trainer.py, placed there by `__graft_entry__.build()` in the build container (where /root/reference is
mounted).  The directory is git-ignored (it is not product source) but travels to the GPU box with the snapshot.  Nothing here
edits those files: missing third-party modules (matplotlib, asteroid, tensorboardX, pesq, pystoi, ...) are stubbed in
sys.modules BEFORE the import, and two configuration values are set the way a user edits config.py (DEVICE = 'cpu',
window = 'hann': scipy >= 1.13 dropped the 'hanning' alias of the same periodic window).
"""
import contextlib
import importlib
import io
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
FILES = ("config.py", "models.py", "tools_for_model.py", "tools_for_loss.py", "trainer.py")
_SHADOWED = ("config", "models", "tools_for_model", "tools_for_loss", "trainer")


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def copy_from(src="/root/reference"):
    """Build-container step: byte-identical copies into the git-ignored baseline/_ref/."""
    import shutil
    if not os.path.isdir(src):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    return True


class _Stub:
    def __init__(self, *a, **k):
        pass

    def to(self, *a, **k):
        return self


def _stub_third_party():
    for n in ["matplotlib", "matplotlib.pylab", "asteroid", "asteroid.losses", "asteroid_filterbanks", "tensorboardX", "pesq", "pystoi",
              "oct2py", "librosa", "soundfile"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["asteroid.losses"].SingleSrcPMSQE = _Stub
    sys.modules["asteroid.losses"].PITLossWrapper = _Stub
    sys.modules["asteroid_filterbanks"].STFTFB = _Stub
    sys.modules["asteroid_filterbanks"].Encoder = _Stub
    sys.modules["asteroid_filterbanks"].transforms = None
    sys.modules["tensorboardX"].SummaryWriter = _Stub
    sys.modules["pesq"].pesq = lambda *a, **k: 0.0
    sys.modules["pystoi"].stoi = lambda *a, **k: 0.0


class Reference:
    """The reference's modules, imported from baseline/_ref with the repo's drop-in modules of the same names hidden."""

    def __init__(self, model):
        self.model_name = model
        saved = {n: sys.modules.pop(n) for n in _SHADOWED if n in sys.modules}
        path = list(sys.path)
        try:
            _stub_third_party()
            sys.path.insert(0, REF_DIR)
            with contextlib.redirect_stdout(io.StringIO()):
                cfg = importlib.import_module("config")          # prints a banner (config.py:94-107)
            cfg.DEVICE, cfg.window = "cpu", "hann"               # BEFORE importing models (models.py:24 freezes win_type)
            cfg.lstm, cfg.skip_type, cfg.perceptual = "complex", True, False
            cfg.loss = "MSE" if model == "fullsubnet" else "SI-SNR"
            cfg.masking_mode = "C"
            self.cfg = cfg
            self.models = importlib.import_module("models")
            self.tools = importlib.import_module("tools_for_model")
        finally:
            sys.path[:] = path
            self._mods = {n: sys.modules.pop(n) for n in _SHADOWED if n in sys.modules}
            sys.modules.update(saved)

    def make_step(self, B, L=48000):
        """One train step = the loop body of trainer.model_train (trainer.py:27-37) / fullsubnet_train (:97-112)."""
        import torch
        torch.manual_seed(0)
        g = torch.Generator().manual_seed(1234)
        noisy = (torch.rand(B, L, generator=g) * 2 - 1) * 0.1
        clean = (torch.rand(B, L, generator=g) * 2 - 1) * 0.1
        if self.model_name == "fullsubnet":
            model = self.models.FullSubNet().train()
        else:
            model = self.models.DCCRN(masking_mode="C").train()
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        tools = self.tools

        def step():
            inputs, targets = noisy.float(), clean.float()
            if self.model_name == "fullsubnet":
                noisy_complex = tools.stft(inputs)
                clean_complex = tools.stft(targets)
                noisy_mag, _ = tools.mag_phase(noisy_complex)
                cirm = tools.build_complex_ideal_ratio_mask(noisy_complex, clean_complex)
                crm = model(noisy_mag)
                loss = model.loss(cirm, crm)
            else:
                _, _, outputs = model(inputs, targets)
                loss = model.loss(outputs, targets)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return float(loss.detach())
        return step


def load(model):
    """Reference(...) or None when baseline/_ref has not been populated (then the oracle port is timed instead)."""
    if not available():
        return None
    try:
        return Reference(model)
    except Exception as e:                      # a broken copy must not take the GPU measurement down with it
        sys.stderr.write(f"refshim: the reference in {REF_DIR} could not be imported ({type(e).__name__}: {e}); using the oracle port\n")
        return None


def load_trainer():
    """The UNMODIFIED reference trainer.py (baseline/_ref) imported against whatever `models` / `tools_for_model` modules are on
    sys.path - i.e. the repo's drop-ins: the boundary test drives trainer.model_train / model_perceptual_train /
    fullsubnet_train exactly as train_interface.py:63-77 would.  tools_for_estimate (PESQ.so / pystoi / oct2py, validation
    only) is stubbed.  Returns None when baseline/_ref has not been populated."""
    path = os.path.join(REF_DIR, "trainer.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    est = types.ModuleType("tools_for_estimate")
    est.cal_pesq = lambda *a, **k: 0.0
    est.cal_stoi = lambda *a, **k: 0.0
    saved = sys.modules.get("tools_for_estimate")
    sys.modules["tools_for_estimate"] = est
    try:
        spec = importlib.util.spec_from_file_location("sefd_reference_trainer", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        if saved is None:
            sys.modules.pop("tools_for_estimate", None)
        else:
            sys.modules["tools_for_estimate"] = saved
    return mod
