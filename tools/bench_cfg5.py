"""cfg5 geometry timing (mono 192 kHz, FFT 65536, hop 1024, log max-pool to 1080 rows): frames/s of the device path."""
import sys, pathlib, time
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import torch
from jadespectrogram_b200 import Engine
fs, N, hop, R = 192000.0, 65536, 1024, 1080
n = int(fs) * 600
eng = Engine(0, sample_rate=fs, fft_size=N, hop=hop, channels=1, row_map="log_maxpool", rows=R, fmin=20.0, fmax=96000.0)
ncols = eng.columns_for(n)
d_in = torch.empty((1, 1, n), dtype=torch.float32, device="cuda")
eng.synth_device(d_in.data_ptr(), 1, 1, n, n, n, kind="mix", seed=5)
d_pix = torch.empty((ncols, R), dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.render_device(d_in.data_ptr(), 1, n, n, n, 0, ncols, d_pix.data_ptr(), None)
eng.sync()
t = []
for _ in range(5):
    eng.render_device(d_in.data_ptr(), 1, n, n, n, 0, ncols, d_pix.data_ptr(), None)
    eng.sync()
    t.append(eng.last_kernel_seconds())
s = sorted(t)[len(t) // 2]
print(f"cfg5 pooled: kernel {eng.kernel_name} {ncols} frames in {s*1e3:.2f} ms -> {ncols/s/1e6:.3f} M frames/s, "
      f"{ncols/s*8416/1e9:.1f} GB/s algorithmic, {ncols/s*2.62e6/1e12:.2f} TFLOP/s (2.5 N log2 N)")
