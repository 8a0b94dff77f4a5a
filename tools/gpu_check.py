"""Diagnostic run for the GPU box: prints max errors of every CUDA op against the oracle, one
section at a time (a failing section does not stop the next unless the CUDA context died)."""
import os
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402
import nerf_keras_b200 as nk  # noqa: E402
from nerf_keras_b200 import _lib  # noqa: E402
from tests.util import golden_weights, load_golden  # noqa: E402


def section(name, fn):
    t0 = time.time()
    try:
        fn()
        torch.cuda.synchronize()
        print(f"[ok]   {name}  ({time.time() - t0:.2f}s)", flush=True)
    except Exception:
        print(f"[FAIL] {name}", flush=True)
        traceback.print_exc()
        sys.stdout.flush()


def rays():
    pose = O.pose_spherical(-63.0, -41.0, 4.0)
    o_ref, d_ref = O.get_rays(800, 800, 1111.1111, pose)
    o, d = nk.get_rays(800, 800, 1111.1111, pose)
    print("   rays bit-exact:", np.array_equal(o.cpu().numpy(), o_ref.numpy()), np.array_equal(d.cpu().numpy(), d_ref.numpy()),
          "max|dd|", float((d.cpu() - d_ref).abs().max()))
    u = np.random.default_rng(3).random(64, dtype=np.float32)
    t_ref = O.generate_t_vals(2.0, 6.0, 4096, 64, True, u=u)
    t = nk.generate_t_vals(2.0, 6.0, 4096, 64, True, u=u)
    print("   t bit-exact:", np.array_equal(t.cpu().numpy(), t_ref.numpy()), float((t.cpu() - t_ref).abs().max()))


def gemm():
    for mode, N, K in [(0, 128, 64), (0, 128, 256), (0, 256, 128), (1, 128, 16), (1, 128, 128), (1, 256, 128)]:
        gen = torch.Generator().manual_seed(1)
        if mode == 0:
            a = torch.randn(128, K, generator=gen); b = torch.randn(N, K, generator=gen)
            ref = a.bfloat16().float() @ b.bfloat16().float().T
        else:
            a = torch.randn(K, 128, generator=gen); b = torch.randn(K, N, generator=gen)
            ref = a.bfloat16().float().T @ b.bfloat16().float()
        c = torch.zeros(128, N, device="cuda")
        a_d, b_d = a.cuda(), b.cuda()
        _lib.check(_lib.lib().nerf_selftest_gemm(a_d.data_ptr(), b_d.data_ptr(), c.data_ptr(), 128, N, K, mode,
                                                 torch.cuda.current_stream().cuda_stream), "selftest")
        torch.cuda.synchronize()
        err = (c.cpu() - ref).abs().max().item()
        print(f"   selftest mode={mode} N={N} K={K}: max err {err:.3e}  (ref max {ref.abs().max().item():.2f})", flush=True)


def trainer(g, wc, wf, precision, batch=None):
    mc = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mf = nk.create_nerf_complete_model(8, 256, 4, 10, 4)
    mc.set_flat_weights(O.flatten_weights(wc)); mf.set_flat_weights(O.flatten_weights(wf))
    tr = nk.NeRFTrainer(mc, mf, batch or g["o"].shape[0], int(g["Nc"]), int(g["Nf"]), 10, 4, precision=precision)
    tr.build()
    return tr


def fwd(precision, label):
    def run():
        for name in ("lego_small", "fern_small"):
            g = load_golden(name)
            wc, wf = golden_weights(g)
            tr = trainer(g, wc, wf, precision)
            rgbs, depths, ws, preds = tr.forward_pass(g["o"], g["d"], g["t"], 10, 4, u_pdf=g["u_pdf"])
            torch.cuda.synchronize()
            print(f"   {label} {name}: rgb_c {np.abs(rgbs[0].cpu().numpy() - g['rgb_c']).max():.3e} "
                  f"rgb_f {np.abs(rgbs[1].cpu().numpy() - g['rgb_f']).max():.3e} "
                  f"pred_c {np.abs(preds[0].cpu().numpy() - g['pred_c']).max():.3e} "
                  f"pred_f {np.abs(preds[1].cpu().numpy() - g['pred_f']).max():.3e} "
                  f"w_c {np.abs(ws[0].cpu().numpy() - g['wt_c']).max():.3e}", flush=True)
    return run


def speed():
    g = load_golden("fern_small")
    wc, wf = golden_weights(g)
    B = 4096
    pose = O.pose_spherical(10.0, -30.0, 4.0)
    o, d = nk.get_rays(80, 80, 100.0, pose)
    o, d = o.reshape(-1, 3)[:B].contiguous(), d.reshape(-1, 3)[:B].contiguous()
    t = nk.generate_t_vals(2.0, 6.0, B, 64, True, u=g["u_t"])
    g2 = dict(g); g2["o"] = o
    tr = trainer(g2, wc, wf, nk.PRECISION_BF16_TC, batch=B)
    u = torch.rand(B, 128, device="cuda")
    for _ in range(3):
        tr.forward_pass(o, d, t, u_pdf=u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 10
    for _ in range(n):
        tr.forward_pass(o, d, t, u_pdf=u)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    flops = B * 256 * 1186816
    print(f"   forward_pass tcgen05 B={B}: {ms:.3f} ms/step  {B / ms * 1e3:.3e} rays/s  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    section("rays/tvals", rays)
    section("forward fp32", fwd(nk.PRECISION_FP32, "fp32"))
    section("tcgen05 selftest gemm", gemm)
    section("forward tcgen05", fwd(nk.PRECISION_BF16_TC, "tc"))
    section("speed", speed)
