"""Does a replay call ever track with the PREVIOUS call's feature lists?  Two feature sets are alternated between calls
on one context; every result must equal what a fresh context computes for the same set."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
sfe = importlib.import_module("slam-robot_b200"); synth = importlib.import_module("slam-robot_b200.synth")
B = 512
dev = torch.device("cuda", 0)
seq = synth.make_sequence(77, B + 2, bench.H, bench.W, stride=2, device=dev).cpu().pin_memory()
sets = [np.concatenate([synth.make_features(1000 * s + p, bench.NFEAT, bench.H, bench.W, margin=16) for p in range(B)]).astype(np.float32) for s in (1, 2)]
n = B * bench.NFEAT
kw = dict(depth=bench.DEPTH, levels=bench.LEVELS, thr=bench.THR, maxit=bench.MAXIT, fb_max=bench.FB_MAX, n_per_pair=bench.NFEAT)
def run(fe, pts):
    h = fe.pinned((n, 2), np.float32); h[...] = pts
    r = fe.replay_sequence(seq, 2, h, h, **kw)
    return {k: np.array(v) for k, v in r.items()}
ref = []
for s in range(2):
    fe = sfe.FrontEnd(0); ref.append(run(fe, sets[s])); fe.close()
fe = sfe.FrontEnd(0)
bad = 0
for it in range(6):
    s = it & 1
    r = run(fe, sets[s])
    same = all(np.array_equal(r[k].view(np.uint8), ref[s][k].view(np.uint8)) for k in ref[s])
    print("call %d (set %d): %s" % (it, s, "identical to a fresh context" if same else "DIFFERS"), flush=True)
    bad += not same
sys.exit(1 if bad else 0)
