#!/usr/bin/env python
"""Runs the UMMA descriptor laboratory (csrc/umma_lab.cu) over the addressing modes the tensor-core kernels
rely on and prints, per case, which address model the hardware followed.  GPU only.

    python tools/umma_lab_report.py > gpurun_out/umma_lab.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "torch-mednet_b200"))
from mednet_b200 import _abi  # noqa: E402
from mednet_b200._abi import check, lib, make  # noqa: E402


def encode(vals, fmt):
    t = torch.from_numpy(vals.astype(np.float32))
    t = t.to(torch.float16 if fmt == 0 else torch.bfloat16)
    return t.view(torch.int16).numpy().view(np.uint16)


def operand(vals_flat, rb, n_mn, ksteps, off, lbo, sbo, mn, kstep, swap=False):
    """Logical operand [n_mn][16*ksteps] read from the un-swizzled image under the canonical-layout model."""
    if swap:
        lbo, sbo = sbo, lbo
    atom = rb // 2
    i = np.arange(n_mn)[:, None]
    k = np.arange(16 * ksteps)[None, :]
    s, kk = k // 16, k % 16
    if mn:
        addr = off + s * kstep + (i % atom) * 2 + (i // atom) * lbo + (kk % 8) * rb + (kk // 8) * sbo
    else:
        addr = off + s * kstep + (i % 8) * rb + (i // 8) * sbo + kk * 2
    return vals_flat[addr // 2]


def run_case(name, rb, rows, M, N, ksteps, a, b, a_fmt=1, b_fmt=1, seed=0):
    rng = np.random.default_rng(seed)
    elems = rb // 2
    vals = rng.integers(-4, 5, size=(rows, elems)).astype(np.float32)
    half = rows // 2
    img = np.empty((rows, elems), np.uint16)
    img[:half] = encode(vals[:half], a_fmt)      # A operands live in the first half of the image,
    img[half:] = encode(vals[half:], b_fmt)      # B operands in the second half
    try:
        g = torch.from_numpy(img.view(np.int16)).cuda()
        out = torch.full((128, N), -777.0, dtype=torch.float32, device="cuda")
    except Exception as e:
        return {"case": name, "verdict": "error: " + str(e)[:120]}
    p = make("mednet_umma_lab_params", g=g.data_ptr(), out=out.data_ptr(), rows=rows, row_bytes=rb, M=M, N=N, ksteps=ksteps,
             a_off=a["off"], a_lbo=a["lbo"], a_sbo=a["sbo"], a_mn_major=a["mn"], a_kstep=a["kstep"],
             b_off=b["off"], b_lbo=b["lbo"], b_sbo=b["sbo"], b_mn_major=b["mn"], b_kstep=b["kstep"], a_fmt=a_fmt, b_fmt=b_fmt)
    try:
        check(lib().mednet_umma_lab(_abi.C.byref(p), torch.cuda.current_stream().cuda_stream), "umma_lab")
        torch.cuda.synchronize()
    except Exception as e:                       # a faulting descriptor poisons the context: report and go on
        return {"case": name, "verdict": "error: " + str(e)[:120]}
    got = out.cpu().numpy()[:M]
    flat = vals.reshape(-1)
    verdict = "none"
    for sa in (False, True):
        for sb in (False, True):
            A = operand(flat, rb, M, ksteps, swap=sa, **a)
            B = operand(flat, rb, N, ksteps, swap=sb, **b)
            if np.array_equal(A @ B.T, got):
                verdict = f"match(a_swapped={sa},b_swapped={sb})"
                break
        if verdict != "none":
            break
    return {"case": name, "verdict": verdict}


def cases_for(rb, R=1024):
    H = R // 2
    hb = H * rb                      # byte offset of the B half
    atom = rb // 2
    kb = rb // 32                    # K-major: 16-element steps per row
    km = dict(off=0, lbo=16, sbo=8 * rb, mn=0, kstep=32)
    nblk = 128 // atom
    la = hb // nblk                  # A atoms spread over the first half of the image
    lb = hb // (256 // atom)         # B atoms (N = 256) spread over the second half
    mnA = dict(off=0, lbo=la, sbo=8 * rb, mn=1, kstep=16 * rb)
    c = []
    c.append((f"rb{rb} kmajor A + kmajor B (sanity)", rb, R, 128, 64, kb, km, dict(off=hb, lbo=16, sbo=8 * rb, mn=0, kstep=32), 1, 1))
    c.append((f"rb{rb} kmajor A + MN-major B N=atom, 1 step", rb, R, 128, atom, 1, km,
              dict(off=hb, lbo=16, sbo=8 * rb, mn=1, kstep=16 * rb), 1, 1))
    c.append((f"rb{rb} kmajor A + MN-major B N=atom, {kb} steps", rb, R, 128, atom, kb, km,
              dict(off=hb, lbo=16, sbo=8 * rb, mn=1, kstep=16 * rb), 1, 1))
    c.append((f"rb{rb} kmajor A + MN-major B N=2 atoms lbo=4096", rb, R, 128, 2 * atom, 1, km,
              dict(off=hb, lbo=4096, sbo=8 * rb, mn=1, kstep=16 * rb), 1, 1))
    c.append((f"rb{rb} kmajor A + MN-major B N=3 atoms chained lbo=rb, pitch-10 sbo, start +3 rows", rb, R, 128, 3 * atom, kb, km,
              dict(off=hb + 3 * rb, lbo=rb, sbo=10 * rb, mn=1, kstep=20 * rb), 1, 1))
    c.append((f"rb{rb} kmajor A + MN-major B N=3 atoms chained, start +11 rows", rb, R, 128, 3 * atom, 1, km,
              dict(off=hb + 11 * rb, lbo=rb, sbo=10 * rb, mn=1, kstep=20 * rb), 1, 1))
    c.append((f"rb{rb} MN-major A M=128 ({nblk} atoms, lbo={la}) + MN-major B N=atom", rb, R, 128, atom, 2, mnA,
              dict(off=hb, lbo=16, sbo=8 * rb, mn=1, kstep=16 * rb), 1, 1))
    c.append((f"rb{rb} MN-major A M=128 chained lbo=rb + MN-major B chained N=3 atoms", rb, R, 128, 3 * atom, 2,
              dict(off=5 * rb, lbo=rb, sbo=8 * rb, mn=1, kstep=16 * rb),
              dict(off=hb + 7 * rb, lbo=rb, sbo=10 * rb, mn=1, kstep=20 * rb), 1, 1))
    c.append((f"rb{rb} MN-major A (f16) + MN-major B (bf16) mixed formats", rb, R, 128, atom, 2, mnA,
              dict(off=hb, lbo=16, sbo=8 * rb, mn=1, kstep=16 * rb), 0, 1))
    c.append((f"rb{rb} MN-major A M=128 + MN-major B N=256 (lbo={lb})", rb, R, 128, 256, 1, mnA,
              dict(off=hb, lbo=lb, sbo=8 * rb, mn=1, kstep=16 * rb), 1, 1))
    return c


def timing_cases():
    """SS-mode issue cost per MMA for several N (K-major operands, 128-byte rows): cycles / MMA."""
    rb, R = 128, 1024
    hb = 512 * rb
    out = []
    for n in (16, 32, 48, 64, 96, 128, 192, 256):
        out.append((n, dict(off=0, lbo=16, sbo=1024, mn=0, kstep=32), dict(off=hb, lbo=16, sbo=1024, mn=0, kstep=32)))
    return out


def run_timing(n, a, b, iters=256, mn=False, nacc=1, M=128):
    rb, R = 128, 1024
    g = torch.zeros((R, rb // 2), dtype=torch.int16, device="cuda")
    out = torch.zeros((128, n), dtype=torch.float32, device="cuda")
    cyc = torch.zeros(4, dtype=torch.int64, device="cuda")
    if mn:
        a = dict(off=0, lbo=32768, sbo=1024, mn=1, kstep=2048)
        b = dict(off=512 * rb, lbo=128 if n > 64 else 16, sbo=1280, mn=1, kstep=2560)
    p = make("mednet_umma_lab_params", g=g.data_ptr(), out=out.data_ptr(), rows=R, row_bytes=rb, M=M, N=n, ksteps=4,
             a_off=a["off"], a_lbo=a["lbo"], a_sbo=a["sbo"], a_mn_major=a["mn"], a_kstep=a["kstep"],
             b_off=b["off"], b_lbo=b["lbo"], b_sbo=b["sbo"], b_mn_major=b["mn"], b_kstep=b["kstep"], a_fmt=1, b_fmt=1,
             iters=iters, cycles=cyc.data_ptr(), nacc=nacc)
    check(lib().mednet_umma_lab(_abi.C.byref(p), torch.cuda.current_stream().cuda_stream), "umma_lab")
    torch.cuda.synchronize()
    c = cyc.cpu().tolist()
    if nacc == -3:                       # two issuers: wall clocks from first issue to the later completion, per MMA of EACH issuer
        return (max(c[1], c[3]) - c[2]) / (iters * 4)
    return c[0] / (iters * 4)


def timing_only():
    _, a, b = timing_cases()[0]
    for M in (128, 64):
        for n in (16, 48, 64, 96, 128, 192, 256):
            for nacc in (-1, -3):
                print(json.dumps({"timing": "lean K-major", "M": M, "N": n, "issuers": 2 if nacc == -3 else 1,
                                  "cycles_per_mma_per_issuer": run_timing(n, a, b, nacc=nacc, M=M)}), flush=True)


def child(rb, start):
    cs = cases_for(rb)
    for i in range(start, len(cs)):
        name, rbb, R, M, N, ks, a, b, af, bf = cs[i]
        r = run_case(name, rbb, R, M, N, ks, a, b, a_fmt=af, b_fmt=bf)
        r["index"] = i
        print(json.dumps(r), flush=True)
        if r["verdict"].startswith("error"):
            sys.exit(3)
    if rb == 128:
        for n, a, b in timing_cases():
            print(json.dumps({"timing": "K-major SS", "N": n, "cycles_per_mma": run_timing(n, a, b)}), flush=True)
            if n in (48, 64, 192):
                print(json.dumps({"timing": "MN-major SS (chained B)", "N": n, "cycles_per_mma": run_timing(n, a, b, mn=True)}),
                      flush=True)


def main():
    import subprocess
    if len(sys.argv) >= 2 and sys.argv[1] == "--timing":
        timing_only()
        return
    if len(sys.argv) >= 4 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        return
    for rb in ([int(a) for a in sys.argv[1:]] or [128, 64, 32]):
        start, n = 0, len(cases_for(rb))
        while start < n:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", str(rb), str(start)],
                               capture_output=True, text=True, timeout=300)
            lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
            for l in lines:
                print(l, flush=True)
            done = [json.loads(l) for l in lines if "index" in l]
            if r.returncode == 0:
                break
            start = (done[-1]["index"] + 1) if done else start + 1
            if not done:
                print(json.dumps({"case": f"rb{rb} #{start - 1}", "verdict": "crash: " + r.stderr[-200:]}), flush=True)


if __name__ == "__main__":
    main()
