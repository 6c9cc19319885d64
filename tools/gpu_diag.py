"""Staged on-GPU bring-up diagnostics (developer tool, not part of the product or the test suite).

    python tools/gpu_diag.py [stage ...]      # each stage runs in its own subprocess under a timeout

Compares the CUDA path with torch ops on the same device (quick, diagnostic only; the real parity tests
in tests/ use the CPU oracle).
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-misinformation-detection_b200"))


def _ref_scores(q, c, eps=1e-12, operand=None):
    import torch
    qn = q.float() / q.float().norm(dim=1, keepdim=True).clamp_min(eps)
    cn = c.float() / c.float().norm(dim=1, keepdim=True).clamp_min(eps)
    if operand == "bf16":
        qn, cn = qn.bfloat16().float(), cn.bfloat16().float()
    elif operand == "fp16":
        qn, cn = qn.half().float(), cn.half().float()
    elif operand == "fp8":
        qn, cn = (qn * 256).to(torch.float8_e4m3fn).float() / 256, (cn * 256).to(torch.float8_e4m3fn).float() / 256
    return (qn.double() @ cn.double().T)


def stage_k1():
    import torch
    import mmd_retrieval as m
    from mmd_retrieval import _lib
    torch.manual_seed(0)
    for dim in (768, 2048, 100, 4100):
        for sdt in (torch.float32, torch.float16, torch.bfloat16):
            x = torch.randn(1000, dim, device="cuda").to(sdt)
            x[5] = 0
            rows, inv = m.normalize_cast(x, "bf16", _lib.SIDE_CORPUS, True, 1e-12)
            kd, rb = m.ops.prepared_layout("bf16", dim)
            got = rows.view(torch.bfloat16).float()[:, :dim]
            ref = (x.float() / x.float().norm(dim=1, keepdim=True).clamp_min(1e-12)).bfloat16().float()
            pad = rows.view(torch.bfloat16).float()[:, dim:]
            print(f"k1 dim={dim} src={sdt} maxdiff={(got-ref).abs().max().item():.3e} mismatches={(got!=ref).sum().item()} "
                  f"pad_nonzero={(pad!=0).sum().item()} inv_err={(inv - 1/x.float().norm(dim=1).clamp_min(1e-12)).abs().max().item():.3e}")
    x = torch.randn(64, 768, device="cuda")
    rows, _ = m.normalize_cast(x, "fp32", _lib.SIDE_QUERY, True, 1e-12)
    limbs = rows.view(torch.bfloat16).float().view(64, 6, 768)
    xn = x / x.norm(dim=1, keepdim=True)
    # query side limb order: 3,1,2,2,1,1 -> limbs[1]+limbs[2]+limbs[0] reconstructs x
    rec = limbs[:, 1] + limbs[:, 2] + limbs[:, 0]
    print("k1 split reconstruct err", (rec - xn).abs().max().item())
    rows8, _ = m.normalize_cast(x, "fp8", _lib.SIDE_CORPUS, True, 1e-12)
    got8 = rows8.view(torch.float8_e4m3fn).float() / 256
    ref8 = (xn * 256).to(torch.float8_e4m3fn).float() / 256
    print("k1 fp8 mismatches", (got8 != ref8).sum().item(), "of", got8.numel())


def _dense_case(Q, N, D, op, seed=0):
    import torch
    import mmd_retrieval as m
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(Q, D, device="cuda", generator=g)
    c = torch.randn(N, D, device="cuda", generator=g)
    got = m.dense_scores(q, c, metric="cos", dtype=op)
    torch.cuda.synchronize()
    ref = _ref_scores(q, c, operand=None if op == "fp32" else op)
    err = (got.double() - ref).abs()
    bad = (err > 1e-4).nonzero()
    print(f"dense Q={Q} N={N} D={D} op={op}: maxerr={err.max().item():.3e} bad={bad.shape[0]}", end="")
    if bad.shape[0]:
        print(" first bad", bad[:6].tolist(), "got", got[bad[0, 0], bad[0, 1]].item(), "ref", ref[bad[0, 0], bad[0, 1]].item())
        rows_bad = torch.unique(bad[:, 0])[:16].tolist()
        cols_bad = torch.unique(bad[:, 1])[:16].tolist()
        print("   bad rows", rows_bad, "bad cols", cols_bad)
    else:
        print()
    return err.max().item()


def stage_dense1():
    _dense_case(128, 256, 64, "bf16")


def stage_dense2():
    for (Q, N, D) in [(128, 256, 128), (128, 256, 768), (256, 512, 64), (300, 1000, 200), (1000, 5000, 768), (37, 41, 2048)]:
        _dense_case(Q, N, D, "bf16")
    _dense_case(200, 700, 768, "fp16")
    _dense_case(200, 700, 768, "fp8")
    _dense_case(200, 700, 768, "fp32")


def _topk_case(Q, N, D, k, op, rescore, seed=0):
    import torch
    import mmd_retrieval as m
    g = torch.Generator(device="cuda").manual_seed(seed)
    q = torch.randn(Q, D, device="cuda", generator=g)
    c = torch.randn(N, D, device="cuda", generator=g)
    pc = m.prepare_corpus(c, dtype=op)
    s, i = m.topk(q, pc, k, rescore_exact=rescore)
    torch.cuda.synchronize()
    ref = _ref_scores(q, c, operand=None if (rescore or op == "fp32") else op)
    rv, ri = torch.sort(ref, dim=1, descending=True, stable=True)
    kk = min(k, N)
    same = (ri[:, :kk] == i).all(dim=1).sum().item()
    sameset = sum(set(a) == set(b) for a, b in zip(ri[:, :kk].tolist(), i.tolist()))
    serr = (s.double() - torch.gather(ref, 1, i.clamp_min(0))).abs().max().item()
    print(f"topk Q={Q} N={N} D={D} k={k} op={op} rescore={rescore}: identical_order={same}/{Q} identical_set={sameset}/{Q} score_err={serr:.3e}")


def stage_topk():
    for args in [(128, 256, 64, 5, "bf16", False), (128, 1000, 768, 10, "bf16", False), (300, 5000, 768, 10, "bf16", False),
                 (1000, 10000, 768, 5, "fp32", False), (1000, 10000, 768, 5, "bf16", True), (64, 20000, 768, 100, "bf16", False),
                 (64, 3000, 768, 40, "bf16", False), (5, 7, 768, 10, "bf16", False), (512, 50000, 2048, 10, "bf16", True)]:
        _topk_case(*args)


def stage_perf():
    import torch
    import mmd_retrieval as m
    for (Q, N, D, k) in [(4096, 50000, 2048, 10), (16384, 1000000, 768, 10)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        q = torch.randn(Q, D, device="cuda", generator=g)
        c = torch.randn(N, D, device="cuda", generator=g)
        pc = m.prepare_corpus(c, dtype="bf16")
        for _ in range(2):
            m.topk(q, pc, k)
        torch.cuda.synchronize()
        m.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            m.topk(q, pc, k)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fused = m.profile_collect()
        m.profile_enable(False)
        fl = 2.0 * Q * N * D
        fm = sum(fused) / len(fused)
        print(f"perf Q={Q} N={N} D={D} k={k}: step {ms:.3f} ms ({Q/ms*1e3:.0f} q/s); fused kernel {fm:.3f} ms = {fl/fm/1e9:.0f} TFLOP/s")


def stage_perf2():
    """Other corners: fp8 operands, K = 100, the reference's small-Q calling pattern (HBM-bound), the fp32 configuration."""
    import torch
    import mmd_retrieval as m
    cases = [  # Q, N, D, k, op, rescore
        (16384, 1000000, 768, 10, "fp8", False), (16384, 1000000, 768, 100, "bf16", False),
        (16384, 1000000, 768, 100, "fp8", False), (16384, 1000000, 512, 10, "bf16", False),
        (1, 1000000, 768, 10, "bf16", False), (100, 1000000, 768, 10, "bf16", False), (128, 4000000, 768, 10, "fp8", False),
        (1000, 10000, 768, 5, "fp32", False), (1000, 10000, 768, 5, "bf16", True)]
    for (Q, N, D, k, op, resc) in cases:
        g = torch.Generator(device="cuda").manual_seed(1)
        q = torch.randn(Q, D, device="cuda", generator=g)
        c = torch.randn(N, D, device="cuda", generator=g)
        pc = m.prepare_corpus(c, dtype=op, keep_source=resc)
        del c
        short = Q * N * D < 4e12                      # sub-millisecond launches: warm the clocks up, time many
        for _ in range(60 if short else 2):
            m.topk(q, pc, k, rescore_exact=resc)
        torch.cuda.synchronize()
        m.profile_enable(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 60 if short else 5
        e0.record()
        for _ in range(n):
            m.topk(q, pc, k, rescore_exact=resc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fused = m.profile_collect()
        m.profile_enable(False)
        fm = sum(fused) / len(fused)
        fl = 2.0 * Q * N * D
        gb = pc.rows.numel() / fm / 1e6
        print(f"perf2 Q={Q} N={N} D={D} k={k} op={op}: step {ms:.3f} ms ({Q/ms*1e3:.0f} q/s); fused {fm:.3f} ms = {fl/fm/1e9:.0f} TFLOP/s, "
              f"corpus stream {gb:.0f} GB/s")
        del pc


def stage_k1perf():
    import torch
    import mmd_retrieval as m
    from mmd_retrieval import _lib
    for (rows, dim, sdt, op) in [(1_000_000, 768, torch.float32, "bf16"), (1_000_000, 768, torch.float16, "bf16"),
                                 (200_000, 2048, torch.float32, "bf16"), (1_000_000, 768, torch.float32, "fp8"),
                                 (16384, 768, torch.float32, "bf16"), (300_000, 768, torch.float32, "fp32")]:
        x = torch.randn(rows, dim, device="cuda").to(sdt)
        for _ in range(2):
            m.normalize_cast(x, op, _lib.SIDE_CORPUS, True, 1e-12)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            out, inv = m.normalize_cast(x, op, _lib.SIDE_CORPUS, True, 1e-12)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        by = x.numel() * x.element_size() + out.numel() + inv.numel() * 4
        print(f"k1perf rows={rows} dim={dim} src={sdt} op={op}: {ms:.3f} ms, {by/ms/1e6:.0f} GB/s (incl. allocator)")
        del x, out


STAGES = {"k1perf": stage_k1perf, "k1": stage_k1, "dense1": stage_dense1, "dense2": stage_dense2, "topk": stage_topk, "perf": stage_perf, "perf2": stage_perf2}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    stages = sys.argv[1:] or list(STAGES)
    for st in stages:
        t0 = time.time()
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", st], capture_output=True, text=True, timeout=None
                           if False else 240)
        print(f"===== stage {st}: rc={r.returncode} ({time.time()-t0:.1f}s)")
        print(r.stdout[-6000:])
        if r.returncode != 0:
            print(r.stderr[-3000:])
