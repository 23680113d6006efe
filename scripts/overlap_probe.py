"""Probe: does preprocessing of batch k+1 overlap the ICP kernel of batch k when two engines (two workspaces) alternate on
two streams?  Prints pairs/s for the plain sequence and for the overlapped one."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
import mgicp_b200 as m

pairs_n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
steps = 4
dev = torch.device("cuda", 0)
scans, pairs, inits, truths = bench.make_workload(pairs_n, 0)
eng = [m.Engine(0), m.Engine(0)]
opts = eng[0].make_opts(loss="l1")
flat, off, _ = eng[0].pack_clouds(scans)
B, S = len(pairs), 3
ps, pt = [p[0] for p in pairs], [p[1] for p in pairs]
md = np.broadcast_to(np.asarray(bench.DISTS), (B, S)).copy()
mi = np.full(S, bench.MAX_IT, np.int32)
xyz = torch.from_numpy(flat).to(dev)
T0 = torch.from_numpy(inits.reshape(B, 16).copy()).to(dev)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]

def run(n_batches, overlapped):
    outs = []
    for k in range(n_batches):
        e = k & 1 if overlapped else 0
        with torch.cuda.stream(streams[e]):
            eng[e].preprocess_device(xyz, off, bench.VOXELS, opts)
            outs.append(eng[e].register_device(ps, pt, md, mi, T0, opts))
    torch.cuda.synchronize()
    return outs

for overlapped in (False, True, False, True):
    run(2, overlapped)
    t = time.perf_counter()
    outs = run(2 * steps, overlapped)
    dt = time.perf_counter() - t
    print(f"overlapped={overlapped}: {2 * steps * B / dt:.1f} pairs/s ({1e3 * dt / (2 * steps):.2f} ms per batch)", flush=True)
a, b = outs[0][0].cpu().numpy(), outs[1][0].cpu().numpy()
print("batches identical:", np.array_equal(a, b))
