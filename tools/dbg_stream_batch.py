import sys, pathlib
ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import signals
from jadespectrogram_b200 import Engine
FS = 48000.0
N, hop, ch = 2048, 512, 2
n = 512 * 40
x = signals.streams(1, ch, n, FS, kind="mix")
eng = Engine(0, sample_rate=FS, fft_size=N, hop=hop, channels=ch, ring_columns=64, max_push=512)
print("kernel", eng.kernel_name)
bpix, bdb = eng.render_batch(x, want_db=True)
bpix2, bdb2 = eng.render_batch(x, want_db=True)
print("batch repeat equal:", np.array_equal(bdb, bdb2))
eng.reset()
cols_p, cols_d = [], []
for b in range(40):
    eng.push(x[0][:, b * 512:(b + 1) * 512])
    p, d, first = eng.fetch()
    cols_p.extend(p); cols_d.extend(d)
sd = np.array(cols_d); sp = np.array(cols_p)
print(sd.shape, bdb.shape)
diff = (sd != bdb[0])
print("cols with diffs:", np.nonzero(diff.any(axis=1))[0])
print("n diffs per col:", diff.sum(axis=1))
j = np.nonzero(diff.any(axis=1))[0]
if len(j):
    c = j[0]; k = np.nonzero(diff[c])[0]
    print("col", c, "bins", k[:20], "stream", sd[c, k[:5]], "batch", bdb[0, c, k[:5]])
    print("max abs diff", np.abs(sd - bdb[0]).max())
print("pix equal:", np.array_equal(sp, bpix[0]))
# which one is the odd one out?
xp = np.concatenate([np.zeros((1, ch, 2048), np.float32), x], axis=2)
eng2 = Engine(0, sample_rate=FS, fft_size=N, hop=hop, channels=ch, ring_columns=64, max_push=512)
ipix, idb = eng2.render_batch(xp, first_col=4, ncols=8, want_db=True)   # interior instantiation: col 4+j of xp == col j of x
for jcol in range(6):
    print("col", jcol, "batch==interior-shifted:", np.array_equal(idb[0, jcol], bdb[0, jcol]),
          " stream==interior-shifted:", np.array_equal(idb[0, jcol], sd[jcol]))
g1, d1 = eng2.render_batch(x, first_col=3, ncols=1, want_db=True)
print("batch col3 alone == batch col3:", np.array_equal(d1[0, 0], bdb[0, 3]))
emu = np.load(str(ROOT / "tools" / "tmp_emu_cols.npy"))
for jcol in range(6):
    print("col", jcol, "max|batch-emu|", np.abs(bdb[0, jcol] - emu[jcol]).max(), " max|stream-emu|", np.abs(sd[jcol] - emu[jcol]).max(),
          " n(batch!=emu)", (bdb[0, jcol] != emu[jcol]).sum(), " n(stream!=emu)", (sd[jcol] != emu[jcol]).sum())
