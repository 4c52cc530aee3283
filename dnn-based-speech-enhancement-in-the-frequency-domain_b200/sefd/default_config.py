"""Values of the reference's config.py (config.py:22-81) that this path reads, used only when no `config`
module is importable (e.g. on a box without the reference checkout).  With the reference on sys.path its own
config.py wins and drives this package unchanged."""
model = 'DCCRN'
loss = 'SDR'
perceptual = False
lstm = 'complex'
masking_mode = 'E'
skip_type = True
DEVICE = 'cuda'
max_epochs = 100
learning_rate = 0.001
batch = 10
dccrn_kernel_num = [32, 64, 128, 256, 256, 256]
fs = 16000
win_len = 400
win_inc = 100
fft_len = 512
window = 'hanning'
rnn_layers = 2
rnn_units = 256
