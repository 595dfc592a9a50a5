import sys
sys.path.insert(0, ".")
import numpy as np, torch
from sitator_b200 import synthetic as syn
from tests import _util as U
system, cfg = syn.make_config("llzo")
frames = system.trajectory(96)
eng = U.engine_for(system); eng.set_frames(frames); eng.reset_status()
seen, words, rows = eng.pass_stats_cached(gram_words=True)
g = eng.gram_words_finish(words).cpu().numpy()
eng.reset_status()
seen2, ga, rows2 = eng.pass_stats_cached(gram_words=False)
ga = ga.cpu().numpy()
lv = eng.fill_dense(dtype=torch.float64).cpu().numpy()
want = np.triu(lv.T @ lv)
# exact reference with python fractions-free: float128
want_ld = np.triu((lv.T.astype(np.longdouble) @ lv.astype(np.longdouble)))
for name, m in (("words", g), ("atomic", ga), ("numpy f64", want)):
    nz = want_ld != 0
    rel = np.abs(m[nz] - want_ld[nz]) / want_ld[nz]
    i = np.argmax(rel)
    print(name, "max rel err vs longdouble %.3e" % float(rel.max()), "at value", float(want_ld[nz][i]), "got", float(m[nz][i]))
w = words.cpu().numpy()
r, c = np.unravel_index(np.argmax(np.where(want_ld != 0, np.abs(g - want_ld) / np.where(want_ld != 0, want_ld, 1), 0)), g.shape)
print("worst entry", r, c, "hi", w[r, c], "lo", w[c, r] if r != c else w[eng.L, r], "value", g[r, c], "want", float(want_ld[r, c]))
print("n rows contributing", int(((lv[:, r] != 0) & (lv[:, c] != 0)).sum()))
