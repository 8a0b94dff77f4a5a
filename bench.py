#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 NeRF path (contract in the task statement, section 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|render|ops|sweep] [--impl ours|reference]
                  [--config config/lego_batch_h256.json] [--rays pinhole|ndc] [--no-graph]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

Default workload (BASELINE.json configs[1], config/lego_batch_h256.json): 8x256 NeRF MLP pair, synthetic Lego-shaped
800x800 views, 4096-ray batches, 64 coarse + 128 fine samples per ray.  `--config config/fern_batch_h256.json` is
configs[2] (Fern-shaped 378x504 views; `--rays pinhole` are the reference's rays, `--rays ndc` the NDC extension).
A "step" is one NeRFTrainer.train_step (forward, loss, backward, gradient all-reduce at N>1, Adam) -- or, with --mode
render, one NeRFTrainer.forward_pass -- over one batch of BATCH_SIZE rays per GPU (weak scaling).  At N > 1 the train
line also carries `strong`: the same step with the GLOBAL batch fixed at BATCH_SIZE (BATCH_SIZE / N rays per GPU), which
is how the reference's strategy scope counts its batch (train_tpu_lego.py:127-163).
--mode ops: stand-alone HBM-bound kernels (ray generation, t-values, compositing forward / backward, resampling + merge,
Adam) at sizes larger than L2, against the measured HBM copy bandwidth.  --mode sweep: BASELINE configs[4], ray-batch
sizes 1 Ki .. 1 Mi x (64+0 | 64+64 | 64+128) samples per ray, train and render.

One JSON line on stdout (rank 0).  `value`: rays/s with the batch resident in HBM; `e2e`: the same through the public API
with pinned-host inputs copied H2D and the step's metrics read back D2H every step.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE_FWD = 1186816          # BASELINE.md section 3 (un-padded, both heads, one net)
CPU_SAMPLE_RAYS = 256                  # bounded CPU sample (rays per CPU step)


def load_conf(path):
    from nerf_keras_b200.config import load_config
    return load_config(path if os.path.isabs(path) else os.path.join(ROOT, path))


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the tensor kernels, from the newest committed
    `ncu --set full` summary under profiles/ (tools/ncu_summary.py --json writes it); {} when there is none."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_traffic.json")))
    if not files:
        return {}, None
    try:
        return json.load(open(files[-1])), os.path.relpath(files[-1], ROOT)
    except Exception:
        return {}, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tensor_sustained": d["bf16_tflops_sustained"], "tensor_burst": d["bf16_tflops"], "hbm": d["hbm_gbs"],
                "source": "measured"}
    return {"tensor_sustained": 1400.0, "tensor_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------------
# synthetic Lego-shaped data
# --------------------------------------------------------------------------------------------------
def lego_poses(n_views, pose_fn):
    rng = np.random.default_rng(0)
    th = rng.uniform(-180.0, 180.0, n_views)
    ph = rng.uniform(-90.0, 0.0, n_views)
    return [pose_fn(float(a), float(b), 4.0) for a, b in zip(th, ph)]


def lego_focal(width):
    return float(np.float32(0.5 * width / np.tan(0.5 * 0.6911112)))


def fern_poses(n_views):
    """Forward-facing poses: identity rotation + yaw / pitch <= 10 degrees, xy translation U(-0.3, 0.3)."""
    rng = np.random.default_rng(2)
    out = []
    for _ in range(n_views):
        yaw, pitch = np.deg2rad(rng.uniform(-10.0, 10.0, 2))
        cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
        Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = (Ry @ Rx).astype(np.float32)
        pose[:2, 3] = rng.uniform(-0.3, 0.3, 2).astype(np.float32)
        out.append(pose)
    return out


def scene_of(conf_path, conf, rays_mode):
    """(name, focal, near, far, pose list factory) of the synthetic scene a config file names."""
    base = os.path.basename(conf_path)
    if base.startswith("ref_"):        # config/ref_*.json: the reference's literal values
        base = base[4:]
    if base.startswith("fern"):
        focal = float(np.float32(407.6 * conf["WIDTH"] / 504.0))
        near, far = (0.0, 1.0) if rays_mode == "ndc" else (1.2, 12.0)
        return "fern", focal, near, far
    return "lego", lego_focal(conf["WIDTH"]), 2.0, 6.0


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs: NVML polled every ~2 ms from a thread
    (the timed region can be as short as 25 ms); `nvidia-smi -lms 20` is the fallback when pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, torch_index):
        self.index, self.proc, self.lines, self.samples = torch_index, None, [], []
        self.stop_flag, self.thread, self.nvml, self.handle = threading.Event(), None, None, None

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(self.index)
        try:
            uuid = "GPU-" + str(props.uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        return pynvml, h

    def start(self):
        try:
            self.nvml, self.handle = self._nvml_handle()
            self.max_mhz = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, mask))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.samples:
                return None
            mask = 0
            for _, m in self.samples:
                mask |= m
            return {"sm_mhz": statistics.median(x for x, _ in self.samples), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(nm for bit, nm in self.REASONS if mask & bit), "samples": len(self.samples),
                    "source": "nvml"}
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle (torch-CPU restatement of the reference; TF/Keras cannot be installed here)
# --------------------------------------------------------------------------------------------------
def cpu_arm(mode, conf, steps, warmup, rays):
    import torch
    import oracle as O
    from oracle.models_ref import _params
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Nc, Nf = conf["NS_COARSE"], conf["NS_FINE"]
    H = W = 64  # rays are drawn from a Lego-shaped view; only `rays` of them are used per CPU step
    pose = lego_poses(1, O.pose_spherical)[0]
    o, d = O.get_rays(H, W, lego_focal(W), pose)
    rng = np.random.default_rng(5)
    sel = rng.choice(H * W, rays, replace=False)
    o, d = o.reshape(-1, 3)[sel].contiguous(), d.reshape(-1, 3)[sel].contiguous()
    t = O.generate_t_vals(2.0, 6.0, rays, Nc, True, u=rng.random(Nc, dtype=np.float32))
    u = torch.from_numpy(rng.random((rays, Nf), dtype=np.float32))
    img = torch.from_numpy(rng.random((rays, 3), dtype=np.float32))
    wc, wf = O.init_weights(42), O.init_weights(43)
    opt = O.KerasAdam(_params(wc) + _params(wf), learning_rate=conf["LEARNING_RATE"])

    def step():
        if mode == "train":     # the reference's gradient semantics: no stop-gradient on the fine samples (models.py:166-175)
            O.train_step(wc, wf, opt, img, o, d, t, 10, 4, Nf, u, stop_grad_samples=False)
        else:
            with torch.no_grad():
                O.forward_pass(wc, wf, o, d, t, 10, 4, Nf, u)

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return {"value": rays / sec, "unit": "rays/s", "cores": cores, "kind": "port",
            "sample": f"{rays} rays x ({Nc}+{Nf}) samples per CPU step (the GPU arm runs {conf['BATCH_SIZE']} per step), median of "
                      f"{steps} steps after {warmup} warm-up; torch-CPU fp32 restatement of the reference pinned against "
                      f"the reference's source (TensorFlow itself is not installable here); reference gradient semantics",
            "ms_per_step": sec * 1e3, "rays_per_cpu_step": rays}


# --------------------------------------------------------------------------------------------------
def timed_region(step_fn, steps, flush, world, dev, torch):
    """K steps bracketed by barrier + synchronize, one CUDA-event pair per step (L2 flushed before each step, outside the
    event pair), max over ranks.  Returns total milliseconds."""
    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
    evs = []
    barrier()
    for i in range(steps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_fn(i)
        b.record()
        evs.append((a, b))
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        total_ms = float(tt.item())
    return total_ms


BACKWARD_OVERLAP_SMS = 0      # --backward-overlap: the opt-in schedule of DESIGN.md 4.3 (weight gradient next to the dX chain)


def make_trainer(nk, conf, B, Nc, Nf, train, use_graph=True, seed=42):
    nk.set_random_seed(seed)
    mk = lambda: nk.create_nerf_complete_model(conf["NUM_LAYERS"], conf["HIDDEN_DIM"], conf["SKIP_LAYER"],
                                               conf["L_XYZ"], conf["L_DIR"], bn=conf["BATCH_NORM"])
    trainer = nk.NeRFTrainer(mk(), mk(), B, Nc, Nf, conf["L_XYZ"], conf["L_DIR"], use_cuda_graph=use_graph,
                             backward_overlap_sms=BACKWARD_OVERLAP_SMS)
    if train:
        trainer.compile(nk.Adam(learning_rate=conf["LEARNING_RATE"]), nk.MeanSquaredError())
    else:
        trainer.build()
    return trainer


def make_batches(nk, torch, dev, scene, conf, rays_mode, B, Nc, Nf, R, rank, explicit_u):
    """R distinct ray batches of the synthetic scene, resident in HBM, plus pinned-host copies."""
    name, focal, near, far = scene
    Hh, Ww = conf["HEIGHT"], conf["WIDTH"]
    poses = lego_poses(R, nk.pose_spherical) if name == "lego" else fern_poses(R)
    rng = np.random.default_rng(1000 + rank)
    u_t = np.random.default_rng(3).random(Nc, dtype=np.float32)
    dev_batches, host_batches = [], []
    for r in range(R):
        o_img, d_img = nk.get_rays(Hh, Ww, focal, poses[r])
        n_pix = Hh * Ww
        sel = torch.from_numpy(rng.choice(n_pix, B, replace=B > n_pix)).to(dev)
        o = o_img.reshape(-1, 3)[sel].contiguous()
        d = d_img.reshape(-1, 3)[sel].contiguous()
        if rays_mode == "ndc":
            o, d = nk.ndc_rays(Hh, Ww, focal, 1.0, o, d)
        t = nk.generate_t_vals(near, far, B, Nc, True, u=u_t)
        img = torch.from_numpy(rng.random((B, 3), dtype=np.float32)).to(dev)
        batch = [img, o, d, t]
        if explicit_u and Nf > 0:
            batch.append(torch.from_numpy(rng.random((B, Nf), dtype=np.float32)).to(dev))
        dev_batches.append(tuple(batch))
        host_batches.append(tuple(x.cpu().pin_memory() for x in batch))
        del o_img, d_img
    return dev_batches, host_batches


def bf16_parity(nk, torch, trainer, batch):
    """Measured deviation of the bf16 tensor-core render from the fp32 path of the same library on one benchmark batch
    (the fp32 path is what tests/ pin against the oracle to 1e-5).  UNMASKED: every ray counts.  north_star asks for
    <= 2e-3 per pixel and <= 0.05 dB; the reference's delta = 1e10 on the last sample makes colour discontinuous in the
    last raw sigma at 0, so rays whose last sigma changes sign under bf16 rounding can jump by a whole colour."""
    img, o, d, t = batch[:4]
    B, Nf = o.shape[0], trainer.ns_fine
    u = torch.from_numpy(np.random.default_rng(77).random((B, max(Nf, 1)), dtype=np.float32)).to(o.device)[:, :Nf].contiguous()
    kw = dict(u_pdf=u) if Nf > 0 else {}
    a = trainer.forward_pass(o, d, t, precision=nk.PRECISION_BF16_TC, return_t_all=Nf > 0, **kw)
    b = trainer.forward_pass(o, d, t, precision=nk.PRECISION_FP32, return_t_all=Nf > 0, **kw)
    out = {}
    for name, idx in (("coarse", 0), ("fine", 1)):
        if a[0][idx] is None:
            continue
        e = (a[0][idx] - b[0][idx]).abs().amax(dim=1).double()
        mse = lambda x: float(((x - img) ** 2).mean())
        out[name] = {"max": float(e.max()), "p999": float(torch.quantile(e, 0.999)), "p99": float(torch.quantile(e, 0.99)),
                     "mean": float(e.mean()), "pixels_over_2e-3": int((e > 2e-3).sum()), "pixels": int(e.numel()),
                     "psnr_delta_db": abs(-10 * np.log10(mse(a[0][idx])) + 10 * np.log10(mse(b[0][idx])))}
    out["note"] = ("bf16 tcgen05 path vs the library's fp32 path on one benchmark batch, unmasked; fine pass includes the "
                   "re-drawn sample positions (inverse CDF of bf16 coarse weights)")
    return out


def run_ops(nk, torch, dev, args, peaks):
    """Stand-alone HBM-bound kernels at sizes larger than L2: achieved algorithmic GB/s against the measured copy peak."""
    from nerf_keras_b200 import _lib
    L = _lib.lib()
    st = lambda: torch.cuda.current_stream().cuda_stream
    res = {}

    def bench(name, fn, bytes_per_call, reps=20, trials=5):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        times = []
        for _ in range(trials):            # clocks ramp and settle during the first launches: median of five back-to-back trials
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b) / reps)
        ms = statistics.median(times)
        gbs = bytes_per_call / (ms * 1e-3) / 1e9
        res[name] = {"ms": ms, "ms_best": min(times), "GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm"], "bytes": bytes_per_call}

    H = W = 2000
    pose = nk.pose_spherical(30.0, -30.0, 4.0)
    p12 = (C.c_float * 12)(*[float(v) for v in np.asarray(pose, np.float32)[:3, :4].reshape(-1)])
    o = torch.empty((H, W, 3), device=dev); d = torch.empty_like(o)
    bench("get_rays_2000x2000", lambda: L.nerf_get_rays(H, W, lego_focal(W), p12, o.data_ptr(), d.data_ptr(), st()), H * W * 24)
    B = 1 << 20
    t = torch.empty((B, 64), device=dev)
    bench("generate_t_vals_1Mi_x64", lambda: L.nerf_generate_t_vals(2.0, 6.0, B, 64, 0, 0, t.data_ptr(), st()), B * 64 * 4)
    for N, Bv in ((64, 1 << 20), (192, 1 << 19)):
        preds = torch.randn((Bv, N, 4), device=dev)
        tt = torch.sort(torch.rand((Bv, N), device=dev) * 4 + 2, dim=1).values.contiguous()
        rgb = torch.empty((Bv, 3), device=dev); dep = torch.empty((Bv,), device=dev); w = torch.empty((Bv, N), device=dev)
        acc = torch.empty((Bv,), device=dev)
        bench(f"volume_render_fwd_{N}", lambda: L.nerf_volume_render(preds.data_ptr(), tt.data_ptr(), Bv, N, rgb.data_ptr(),
                                                                     dep.data_ptr(), w.data_ptr(), acc.data_ptr(), st()),
              Bv * (N * 24 + 20))
        drgb = torch.randn((Bv, 3), device=dev); dp = torch.empty_like(preds)
        bench(f"volume_render_bwd_{N}", lambda: L.nerf_volume_render_bwd(preds.data_ptr(), tt.data_ptr(), drgb.data_ptr(), 0, Bv,
                                                                         N, dp.data_ptr(), 0, st()),
              Bv * (N * 36 + 12))
        del preds, tt, w, dp
    Bv = 1 << 18
    tc = torch.sort(torch.rand((Bv, 64), device=dev) * 4 + 2, dim=1).values.contiguous()
    wc = torch.rand((Bv, 64), device=dev); u = torch.rand((Bv, 128), device=dev)
    ta = torch.empty((Bv, 192), device=dev)
    bench("resample_merge_64+128", lambda: L.nerf_resample_merge(tc.data_ptr(), wc.data_ptr(), u.data_ptr(), Bv, 64, 128,
                                                                 ta.data_ptr(), 0, st()), Bv * 1792)
    n = 64 * 1191688
    p_, g_, m_, v_ = (torch.zeros(n, device=dev) for _ in range(4))
    bench("adam_76M_params", lambda: L.nerf_adam_flat(p_.data_ptr(), g_.data_ptr(), m_.data_ptr(), v_.data_ptr(), n, 1, 5e-4,
                                                      1.0, st()), n * 28)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", choices=["train", "render", "ops", "sweep"], default="train")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--config", type=str, default="config/lego_batch_h256.json")
    ap.add_argument("--rays", choices=["pinhole", "ndc"], default="pinhole")
    ap.add_argument("--ns-fine", type=int, default=None, help="override NS_FINE (0 = single-net 64-sample shape)")
    ap.add_argument("--batch", type=int, default=None, help="override BATCH_SIZE (rays per step per GPU)")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel of the step from the host (no CUDA graph)")
    ap.add_argument("--no-overlap", action="store_true", help="one all-reduce after the whole backward (N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--backward-overlap", type=int, default=0,
                    help="experiment (DESIGN.md 4.3): weight-gradient kernel on this many SMs next to the dX chain")
    args = ap.parse_args()
    global BACKWARD_OVERLAP_SMS
    BACKWARD_OVERLAP_SMS = args.backward_overlap
    conf = load_conf(args.config)
    if args.ns_fine is not None:
        conf["NS_FINE"] = args.ns_fine
    if args.batch is not None:
        conf["BATCH_SIZE"] = args.batch
    B, Nc, Nf = conf["BATCH_SIZE"], conf["NS_COARSE"], conf["NS_FINE"]
    spr = Nc + (Nc + Nf if Nf > 0 else 0)            # MLP evaluations per ray
    scene = scene_of(args.config, conf, args.rays)
    metric = {"train": "train_rays_per_sec", "render": "render_rays_per_sec", "ops": "hbm_kernels_frac_of_peak",
              "sweep": "train_rays_per_sec"}[args.mode]
    config = {"workload": f"{args.config}: 8x256 MLP x{2 if Nf else 1} (coarse {Nc}" + (f" + fine {Nf}" if Nf else "") +
                          f" samples/ray), synthetic {scene[0].capitalize()}-shaped {conf['HEIGHT']}x{conf['WIDTH']} views"
                          f" ({args.rays} rays), {B}-ray batch per GPU" +
                          (", BATCH_NORM=true (layer-by-layer path, tcgen05 split-bf16 GEMMs)" if conf.get("BATCH_NORM") else ""),
              "mode": args.mode, "rays_per_step_per_gpu": B, "samples_per_ray": spr, "parallelism": f"dp{args.gpus}",
              "l2": "8 distinct resident ray batches rotated + 256 MiB L2 flush between timed steps",
              "step": "eager launches" if (args.no_graph or conf.get("BATCH_NORM")) else "CUDA-graph replay"}
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        r = cpu_arm("render" if args.mode == "render" else "train", conf, steps, warm, CPU_SAMPLE_RAYS)
        config = dict(config, rays_per_step_per_gpu=CPU_SAMPLE_RAYS,
                      workload=config["workload"] + f" -- CPU arm: a bounded sample of {CPU_SAMPLE_RAYS} rays per step")
        line = {"impl": "reference", "metric": "render_rays_per_sec" if args.mode == "render" else "train_rays_per_sec",
                "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import nerf_keras_b200 as nk
    from nerf_keras_b200 import _lib, models as nkm
    from nerf_keras_b200.dist import init_from_env
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, local, world = init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    L = _lib.lib()
    peaks = measured_peaks()
    steps, warm = args.steps, max(3, args.warmup)

    if args.mode == "ops":
        sampler = ClockSampler(local); sampler.start()
        res = run_ops(nk, torch, dev, args, peaks)
        clocks = sampler.stop()
        worst = min(res, key=lambda k: res[k]["frac_of_hbm_peak"])
        line = {"metric": metric, "value": res[worst]["frac_of_hbm_peak"], "unit": "fraction of measured HBM copy bandwidth "
                "(slowest stand-alone kernel)", "n_gpus": 1, "steps": 100, "warmup": 5, "ms_per_step": res[worst]["ms"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "stand-alone HBM-bound kernels, inputs larger than L2 (126 MB)", "mode": "ops"},
                "ops": res, "clocks": clocks, "gpu_launches": 105 * len(res),
                "roofline": {"bound": "hbm", "kernel": worst, "achieved": res[worst]["GB/s"], "peak": peaks["hbm"], "unit": "GB/s",
                             "frac": res[worst]["frac_of_hbm_peak"], "traffic": None}}
        if rank == 0:
            print(json.dumps(line), flush=True)
        return 0

    if args.mode == "sweep":
        points = []
        for nf in (0, 64, 128):
            for lb in (10, 12, 14, 16, 18, 20):
                Bs = 1 << lb
                cap = min(Bs, 32768 if nf else 131072)            # workspace cap (saved operand images: 10.4 KB per sample, ~90 GB): larger batches run as micro-batches
                c2 = dict(conf, NS_FINE=nf, BATCH_SIZE=Bs)
                tr = make_trainer(nk, c2, cap, Nc, nf, True, use_graph=not args.no_graph)
                db, _ = make_batches(nk, torch, dev, scene, c2, args.rays, Bs, Nc, nf, 2, rank, False)
                fn = lambda i: tr.train_step((db[i % 2][0], db[i % 2][1:4]))
                for i in range(3):
                    fn(i)
                k = max(2, min(steps, (1 << 22) // Bs))
                ms = timed_region(fn, k, None, world, dev, torch) / k
                fr = lambda i: tr.forward_pass(*db[i % 2][1:4], maps_only=True)
                for i in range(2):
                    fr(i)
                rms = timed_region(fr, k, None, world, dev, torch) / k
                per_ray = (Nc + (Nc + nf if nf else 0)) * FLOP_PER_SAMPLE_FWD
                points.append({"rays_per_step_per_gpu": Bs, "ns_coarse": Nc, "ns_fine": nf, "train_ms": ms,
                               "train_rays_per_sec": world * Bs / (ms * 1e-3), "render_ms": rms,
                               "render_rays_per_sec": world * Bs / (rms * 1e-3),
                               "train_frac_of_sustained_bf16": Bs / (ms * 1e-3) * 3 * per_ray / (peaks["tensor_sustained"] * 1e12),
                               "render_frac_of_sustained_bf16": Bs / (rms * 1e-3) * per_ray / (peaks["tensor_sustained"] * 1e12),
                               "micro_batches": Bs // cap})
                tr.release_graphs()
                del tr, db
                torch.cuda.empty_cache()
        best = max(points, key=lambda p: p["train_rays_per_sec"])
        line = {"metric": metric, "value": best["train_rays_per_sec"], "unit": "rays/s (best point of the sweep)",
                "n_gpus": world, "steps": steps, "warmup": 3, "ms_per_step": best["train_ms"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": dict(config, workload="ray-batch sweep 1 Ki .. 1 Mi rays/step x (64+0 | 64+64 | 64+128) samples/ray "
                                                "(BASELINE configs[4]); in-kernel draws; inputs resident"),
                "sweep": points}
        if rank == 0:
            print(json.dumps(line), flush=True)
        from nerf_keras_b200.dist import shutdown
        shutdown()
        return 0

    # ---- model + trainer (random-init weights of the named architecture) -----------------------
    train = args.mode == "train"
    trainer = make_trainer(nk, conf, B, Nc, Nf, train, use_graph=not args.no_graph)
    trainer.overlap_allreduce = not args.no_overlap
    R = 8
    dev_batches, host_batches = make_batches(nk, torch, dev, scene, conf, args.rays, B, Nc, Nf, R, rank, False)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_dev(i):
        img, o, d, t = dev_batches[i % R][:4]
        if train:
            return trainer.train_step((img, (o, d, t)))
        return trainer.forward_pass(o, d, t)[0][1 if Nf else 0]

    rgb_host = torch.empty((B, 3), dtype=torch.float32).pin_memory()
    from nerf_keras_b200.synthetic import HostPrefetcher

    def endless_host_batches():
        j = 0
        while True:
            yield host_batches[j % R]
            j += 1
    prefetcher = HostPrefetcher(endless_host_batches(), dev)     # one instance: its two staging slots keep their addresses
    e2e_state = {"it": None}

    def step_e2e(i):
        # every step's inputs come from pinned host memory; the public HostPrefetcher overlaps the copy of step i+1
        # with step i (the first copy of a run is not overlapped: the iterator is created inside step 0)
        if i == 0 or e2e_state["it"] is None:
            e2e_state["it"] = iter(prefetcher)
        img, o, d, t = next(e2e_state["it"])[:4]
        if train:
            trainer.train_step((img, (o, d, t)))
            return trainer._ctx.metric_sums().tolist()          # one D2H read (16 bytes) of the step's metrics
        rgb = trainer.forward_pass(o, d, t)[0][1 if Nf else 0]
        rgb_host.copy_(rgb, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return rgb_host

    for i in range(max(warm, 2 * R if train and not args.no_graph else warm)):   # every resident batch is seen twice: graphs captured
        step_dev(i)
    for i in range(warm + 2):
        step_e2e(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    total_ms = timed_region(step_dev, steps, flush, world, dev, torch)
    e2e_ms = timed_region(step_e2e, steps, flush, world, dev, torch)
    # per-kernel CUDA-event timers bracket single launches, so this pass enqueues the step eagerly (same kernels, same order)
    launches0 = _lib.launch_count()
    nkm.set_kernel_timing(True)
    eager_ms = timed_region(step_dev, steps, flush, world, dev, torch)
    nkm.set_kernel_timing(False)
    launches_per_step = (_lib.launch_count() - launches0) / steps
    clocks = sampler.stop()               # sampled under load over the timed regions
    kernel_ms, kernel_n = {}, {}
    for kind, nm in ((0, "mlp_fwd"), (1, "mlp_bwd_chain"), (2, "wgrad")):
        a_ms, a_n = C.c_double(), C.c_int64()
        L.nerf_timing_read(kind, C.byref(a_ms), C.byref(a_n))
        if a_n.value:
            kernel_ms[nm], kernel_n[nm] = a_ms.value / steps, a_n.value

    extra = {}
    ms_per_step = total_ms / steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / steps * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in host_batches[0])
    d2h = 16 if train else B * 12
    if train:
        # the metric names render AND train: time forward_pass on the same trainer / batches as well
        def step_render(i, maps_only):
            img, o, d, t = dev_batches[i % R][:4]
            return trainer.forward_pass(o, d, t, maps_only=maps_only)[0][1 if Nf else 0]
        for mo, key in ((False, "render_rays_per_sec"), (True, "render_maps_only_rays_per_sec")):
            for i in range(3):
                step_render(i, mo)
            k = max(3, steps // 2)
            r_ms = timed_region(lambda i: step_render(i, mo), k, flush, world, dev, torch)
            extra[key] = world * B / (r_ms / k * 1e-3)
        extra["render_note"] = ("render_rays_per_sec returns everything the reference's forward_pass returns (rgb, depth, "
                                "weights, raw predictions of both nets); render_maps_only skips the per-sample outputs "
                                "(an extension: what a renderer needs)")
        L.nerf_timing_read(0, C.byref(C.c_double()), C.byref(C.c_int64()))
        extra["parity"] = bf16_parity(nk, torch, trainer, dev_batches[0])
        extra["eager_ms_per_step"] = eager_ms / steps
        extra["glue_ms_per_step"] = ms_per_step - sum(kernel_ms.values())
        per_ray_train = 3 * spr * FLOP_PER_SAMPLE_FWD
        extra["step_frac_of_sustained_bf16"] = (B / (ms_per_step * 1e-3)) * per_ray_train / (peaks["tensor_sustained"] * 1e12)

    # strong scaling (N > 1): the global batch stays at BATCH_SIZE, each GPU takes BATCH_SIZE / N rays
    if train and world > 1 and not args.no_strong and B % world == 0:
        Bl = B // world
        tr_s = make_trainer(nk, conf, Bl, Nc, Nf, True, use_graph=not args.no_graph)
        tr_s.overlap_allreduce = not args.no_overlap
        sb = [tuple(x[rank * Bl:(rank + 1) * Bl].contiguous() for x in b[:4]) for b in dev_batches]
        fn = lambda i: tr_s.train_step((sb[i % R][0], sb[i % R][1:4]))
        for i in range(2 * R):
            fn(i)
        s_ms = timed_region(fn, steps, flush, world, dev, torch) / steps
        extra["strong"] = {"global_batch": B, "rays_per_gpu": Bl, "ms_per_step": s_ms, "rays_per_sec": B / (s_ms * 1e-3),
                           "speedup_vs_weak_step_of_this_run": ms_per_step / s_ms,
                           "efficiency_vs_weak_step_of_this_run": ms_per_step / s_ms / world,
                           "note": "weak step of this run = BATCH_SIZE rays on every GPU (the 1-GPU amount of work + all-reduce)"}
        tr_s.release_graphs()
        del tr_s

    # rooflines of the three tensor kernels of a step (live CUDA-event timers of the eager pass above)
    traffic, traffic_src = ncu_traffic()
    samples_per_step = B * spr
    cands = {}
    if "mlp_fwd" in kernel_ms:
        fl = samples_per_step * FLOP_PER_SAMPLE_FWD
        ach = fl / (kernel_ms["mlp_fwd"] * 1e-3) / 1e12
        cands["roofline_fwd"] = {"bound": "tensor", "kernel": "nerf_mlp_fwd_tc_kernel" + ("<save>" if train else "<render>"),
                                 "achieved": ach, "peak": peaks["tensor_sustained"], "unit": "TFLOP/s",
                                 "frac": ach / peaks["tensor_sustained"],
                                 "traffic": traffic.get("mlp_fwd_train" if train else "mlp_fwd_render"),
                                 "avg_launch_ms": kernel_ms["mlp_fwd"] * steps / kernel_n["mlp_fwd"],
                                 "launches": kernel_n["mlp_fwd"], "share_of_step": kernel_ms["mlp_fwd"] / (eager_ms / steps)}
    if train and "wgrad" in kernel_ms:
        tiles = -(-B * Nc // 128) + (-(-B * (Nc + Nf) // 128) if Nf else 0)
        wg_bytes = tiles * (638976 + 638976)     # every saved activation image + every dZ image of a tile, once
        gbs = wg_bytes / (kernel_ms["wgrad"] * 1e-3) / 1e9
        wg_flop = samples_per_step * FLOP_PER_SAMPLE_FWD        # dW = X^T dZ: the forward's MACs once more
        cands["roofline_wgrad"] = {"bound": "hbm", "kernel": "nerf_wgrad_tc_kernel", "unit": "GB/s", "peak": peaks["hbm"],
                                   "achieved": gbs, "frac": gbs / peaks["hbm"], "traffic": traffic.get("wgrad"),
                                   "tensor_frac_of_sustained": wg_flop / (kernel_ms["wgrad"] * 1e-3) / 1e12 / peaks["tensor_sustained"],
                                   "avg_launch_ms": kernel_ms["wgrad"] * steps / kernel_n["wgrad"], "launches": kernel_n["wgrad"],
                                   "share_of_step": kernel_ms["wgrad"] / (eager_ms / steps)}
        chain_flop = samples_per_step * 2 * (256 * 128 + 8 * 256 * 256)
        ach = chain_flop / (kernel_ms["mlp_bwd_chain"] * 1e-3) / 1e12
        cands["roofline_chain"] = {"bound": "tensor", "kernel": "nerf_mlp_bwd_tc_kernel", "unit": "TFLOP/s",
                                   "peak": peaks["tensor_sustained"], "achieved": ach, "frac": ach / peaks["tensor_sustained"],
                                   "traffic": traffic.get("mlp_bwd_chain"),
                                   "avg_launch_ms": kernel_ms["mlp_bwd_chain"] * steps / kernel_n["mlp_bwd_chain"],
                                   "launches": kernel_n["mlp_bwd_chain"],
                                   "share_of_step": kernel_ms["mlp_bwd_chain"] / (eager_ms / steps)}
    for c in cands.values():
        c["peak_source"] = peaks["source"] + (" (HBM copy)" if c["bound"] == "hbm" else " (sustained bf16)")
        c["traffic_unit"] = f"bytes/launch (ncu --set full, {traffic_src})" if traffic_src else None
    roofline = None
    if cands:   # `roofline` is the kernel with the largest share of the step; the others keep their own keys
        top = max(cands, key=lambda k: cands[k]["share_of_step"])
        roofline = cands.pop(top)
        extra.update(cands)

    line = {"metric": metric, "value": value, "unit": "rays/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "samples_per_sec": value * spr,
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(round(launches_per_step * steps)), "launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "kernel_ms_per_step": kernel_ms}
    line.update(extra)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_arm(args.mode, conf, 3, 1, CPU_SAMPLE_RAYS)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    from nerf_keras_b200.dist import shutdown
    e2e_state["it"] = None
    shutdown(trainer)
    return 0


if __name__ == "__main__":
    sys.exit(main())
