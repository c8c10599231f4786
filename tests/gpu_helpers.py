"""GPU-side helpers of the parity tests (every call goes through the C ABI)."""
import ctypes as C

import numpy as np

from oracle import ofdm_oracle as o


def tx_gpu(eng, pay, pad_for_usrp=False, first_frame=0):
    import torch
    off = np.zeros(len(pay) + 1, dtype=np.int64)
    np.cumsum([len(p) for p in pay], out=off[1:])
    plan = eng.tx_plan(off, pad_for_usrp=pad_for_usrp)
    raw = np.frombuffer(b"".join(pay), dtype=np.uint8).copy() if off[-1] else np.zeros(1, np.uint8)
    d_pay = torch.from_numpy(raw).cuda()
    x = eng.tx_run(plan, d_pay, first_frame=first_frame)
    torch.cuda.synchronize()
    return x, plan


def processed_vectors(vbase, ndata, nvec):
    """Rows of the sampler's vector stream that some sink session equalised on the GPU: the session started
    at frame g covers vectors vbase[g] .. vbase[g] + min(nvec[g], 1 + ndata[g]) of its own frame."""
    rows = []
    for g in range(len(ndata)):
        k = int(min(int(nvec[g]), 1 + int(ndata[g])))
        rows.extend(range(int(vbase[g]), int(vbase[g]) + k))
    return np.array(rows, dtype=np.int64)


def oracle_demapped(r, lay):
    """(vector index, slicer decisions) for every vector the oracle's sink demapped, in order."""
    sink = o.FrameSink(lay)
    out = []
    for v in range(len(r.flags)):
        before = len(sink.sym_log)
        sink.work(r.eq[v], int(r.flags[v]))
        if len(sink.sym_log) > before:
            out.append(v)
    return out
