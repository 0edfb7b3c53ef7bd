"""CPU estimate (through tests/host_shim) of how often the sketch kernel's deferred-atomics branch is taken: hashes every k-mer of a
synthetic 2.5 Mbp genome with the device arithmetic, replays the ULL cell updates in order and groups the k-mers the way a warp
processes them (32 lanes x 16 k-mers per step).  Also evaluates a cheaper per-register THRESHOLD filter (nlz >= top - 2) that was
considered instead of the exact "is this bit new" filter.  Result (p=10, k=16): exact filter -> branch taken in 53 % of the
warp-steps (99 % in the first tenth of the genome, 26 % in the last), 4.6 new bits per taken step; threshold filter -> 88 %.
    python tools/filter_stats.py
"""
import sys, ctypes as C, subprocess, numpy as np
sys.path.insert(0, ".")
from lash_b200 import hostapi
from tools import synth
subprocess.check_call(["g++","-O2","-std=c++17","-shared","-fPIC","-o","/tmp/libdm.so","tests/host_shim/device_math.cpp"])
L = C.CDLL("/tmp/libdm.so")
vp,u64,i32 = C.c_void_p, C.c_uint64, C.c_int
L.dm_kmers.argtypes=[vp,u64,u64,i32,vp]; L.dm_pre.argtypes=[vp,u64,u64,i32,vp,vp]; L.dm_ull_fast.argtypes=[vp,u64,i32,i32,vp,vp,vp]
P=lambda a:a.ctypes.data_as(vp)
p,k = 10,16
seq = synth.genomes(1, 2_500_000, seed=1)[0][0]
packed, nb = hostapi.filter_pack(seq, simd=1)
buf = np.concatenate([packed, np.zeros(16, np.uint8)])
n = len(seq)-k+1
km = np.zeros(n, np.uint64); L.dm_kmers(P(buf), len(packed), len(seq), k, P(km))
g = np.empty_like(km); ghi = np.empty(n, np.uint32); L.dm_pre(P(km), n, 42, 1, P(g), P(ghi))
idx,nlz,rare = (np.empty(n,np.uint32) for _ in range(3)); L.dm_ull_fast(P(ghi), n, p, 1, P(idx), P(nlz), P(rare))
print("rare fraction", rare.mean())
# sequential simulation: event = bit (idx, nlz) not yet seen
key = idx.astype(np.int64)*64 + np.minimum(nlz,63).astype(np.int64)
_, first = np.unique(key, return_index=True)
event = np.zeros(n, bool); event[first] = True
print("events (new bits) total", event.sum(), "rate", event.mean())
# conservative threshold filter: pass if nlz >= top(idx) - 2 at that time -> approximate with running max per register
top = np.full(1<<p, -1, np.int64); passc = np.zeros(n, bool)
# vectorised approximation in blocks of 2048 (registers' tops updated per block)
B=2048
for o in range(0, n, B):
    sl = slice(o, min(n,o+B))
    t = top[idx[sl]]
    passc[sl] = nlz[sl].astype(np.int64) >= t-2
    np.maximum.at(top, idx[sl].astype(np.int64), nlz[sl].astype(np.int64))
print("conservative pass rate", passc.mean())
for name, ev in (("exact filter", event), ("threshold filter", passc)):
    m = n//2048*2048
    e = ev[:m].reshape(-1, 32, 4, 16)         # warp-block, lane, w-step, k-mer
    grp = e.transpose(0,2,1,3).reshape(-1, 512)   # one warp-group = 32 lanes x 16 k-mers of one w-step
    taken = grp.any(axis=1)
    q = [0, len(taken)//10, len(taken)//2, len(taken)-len(taken)//10]
    print(name, ": branch taken in %.3f of warp-groups; events per taken group %.2f; by decile of the genome:" % (taken.mean(), grp.sum()/max(1,taken.sum())),
          [round(float(taken[a:a+len(taken)//10].mean()),3) for a in q])
