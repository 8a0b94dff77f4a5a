#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 NeRF path (contract in the task statement, section 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode train|render] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1, one rank per GPU)

Workload (BASELINE.json configs[1], config/lego_batch_h256.json): 8x256 NeRF MLP pair, synthetic
Lego-shaped 800x800 views, 4096-ray batches, 64 coarse + 128 fine samples per ray.  A "step" is one
NeRFTrainer.train_step (forward, loss, backward, gradient all-reduce at N>1, Adam) -- or, with
--mode render, one NeRFTrainer.forward_pass -- over one batch of 4096 rays per GPU (weak scaling).

One JSON line on stdout (rank 0).  `value`: rays/s with the batch resident in HBM; `e2e`: the same
through the public API with pinned-host inputs copied H2D and the result read back D2H every step.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
# dram__bytes_read.sum + dram__bytes_write.sum of the fused forward kernel per launch (mean of the coarse and fine
# launches of one step) from the committed `ncu --set full` captures: profiles/r1_train_kernels_ncu_summary.txt
# (training variant, writes the saved operand images) and profiles/r1_render_fwd_ncu_summary.txt (render variant)
NCU_TRAFFIC_PER_LAUNCH = {"train": (0.030137e9 + 1.353939e9 + 0.085956e9 + 4.178626e9) / 2,
                          "render": (4.590592e6 + 6.688768e6 + 1.280e3) / 2}
CPU_SAMPLE_RAYS = 256                  # bounded CPU sample (rays per CPU step)


def load_conf():
    from nerf_keras_b200.config import load_config
    return load_config(os.path.join(ROOT, "config", "lego_batch_h256.json"))


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
        if mode == "train":
            O.train_step(wc, wf, opt, img, o, d, t, 10, 4, Nf, u, stop_grad_samples=True)
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
            "sample": f"{rays} rays x ({Nc}+{Nf}) samples per CPU step, median of {steps} steps after {warmup} warm-up; "
                      f"torch-CPU fp32 restatement of the reference (TensorFlow not installable here)",
            "ms_per_step": sec * 1e3}


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", choices=["train", "render"], default="train")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    conf = load_conf()
    B, Nc, Nf = conf["BATCH_SIZE"], conf["NS_COARSE"], conf["NS_FINE"]
    metric = "train_rays_per_sec" if args.mode == "train" else "render_rays_per_sec"
    config = {"workload": "config/lego_batch_h256.json: 8x256 MLP x2 (coarse 64 + fine 128 samples/ray), synthetic "
                          "Lego-shaped 800x800 views, 4096-ray batch per GPU", "mode": args.mode,
              "rays_per_step_per_gpu": B, "samples_per_ray": Nc + Nc + Nf, "parallelism": f"dp{args.gpus}",
              "l2": "8 distinct resident ray batches rotated + 256 MiB L2 flush between timed steps"}
    rank = int(os.environ.get("RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        r = cpu_arm(args.mode, conf, steps, warm, CPU_SAMPLE_RAYS)
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": "rays/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    import torch
    import nerf_keras_b200 as nk
    from nerf_keras_b200 import _lib
    from nerf_keras_b200.dist import init_from_env
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, local, world = init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    dev = torch.device("cuda", local)
    L = _lib.lib()

    # ---- model + trainer (random-init weights of the named architecture) -----------------------
    nk.set_random_seed(42)
    mk = lambda: nk.create_nerf_complete_model(conf["NUM_LAYERS"], conf["HIDDEN_DIM"], conf["SKIP_LAYER"],
                                               conf["L_XYZ"], conf["L_DIR"], bn=conf["BATCH_NORM"])
    coarse, fine = mk(), mk()
    trainer = nk.NeRFTrainer(coarse, fine, B, Nc, Nf, conf["L_XYZ"], conf["L_DIR"])
    if args.mode == "train":
        trainer.compile(nk.Adam(learning_rate=conf["LEARNING_RATE"]), nk.MeanSquaredError())
    else:
        trainer.build()

    # ---- synthetic Lego-shaped batches: R distinct ones, on the device and in pinned host memory --
    R = 8
    Hh, Ww = conf["HEIGHT"], conf["WIDTH"]
    focal = lego_focal(Ww)
    poses = lego_poses(R, nk.pose_spherical)
    rng = np.random.default_rng(1000 + rank)
    u_t = np.random.default_rng(3).random(Nc, dtype=np.float32)
    dev_batches, host_batches = [], []
    for r in range(R):
        o_img, d_img = nk.get_rays(Hh, Ww, focal, poses[r])
        sel = torch.from_numpy(rng.choice(Hh * Ww, B, replace=False)).to(dev)
        o = o_img.reshape(-1, 3)[sel].contiguous()
        d = d_img.reshape(-1, 3)[sel].contiguous()
        t = nk.generate_t_vals(2.0, 6.0, B, Nc, True, u=u_t)
        img = torch.from_numpy(rng.random((B, 3), dtype=np.float32)).to(dev)
        u_pdf = torch.from_numpy(rng.random((B, Nf), dtype=np.float32)).to(dev)
        dev_batches.append((img, o, d, t, u_pdf))
        host_batches.append(tuple(x.cpu().pin_memory() for x in (img, o, d, t, u_pdf)))
        del o_img, d_img
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step_dev(i):
        img, o, d, t, u = dev_batches[i % R]
        if args.mode == "train":
            return trainer.train_step((img, (o, d, t)), u_pdf=u)
        return trainer.forward_pass(o, d, t, u_pdf=u)[0][1]

    rgb_host = torch.empty((B, 3), dtype=torch.float32).pin_memory()

    from nerf_keras_b200.synthetic import HostPrefetcher
    e2e_state = {"it": None}

    def step_e2e(i):
        # every step's inputs come from pinned host memory; the public HostPrefetcher overlaps the copy of step i+1
        # with step i (the first copy of a run is not overlapped: the iterator is created inside step 0)
        if i == 0 or e2e_state["it"] is None:
            e2e_state["it"] = iter(HostPrefetcher((host_batches[j % R] for j in range(1 << 30)), dev))
        img, o, d, t, u = next(e2e_state["it"])
        if args.mode == "train":
            logs = trainer.train_step((img, (o, d, t)), u_pdf=u)
            return torch.stack([logs["loss"], logs["psnr"], logs["loss_coarse"]]).tolist()  # one D2H read of the step's metrics
        rgb = trainer.forward_pass(o, d, t, u_pdf=u)[0][1]
        rgb_host.copy_(rgb, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return rgb_host

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, with_kernel_timing=False):
        evs = []
        barrier()
        if with_kernel_timing:
            L.nerf_timing_enable(1)
        for i in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(i)
            b.record()
            evs.append((a, b))
        barrier()
        if with_kernel_timing:
            L.nerf_timing_enable(0)
        total_ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            total_ms = float(tt.item())
        return total_ms

    for i in range(max(3, args.warmup)):
        step_dev(i)
        step_e2e(i)
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count()
    total_ms = timed(step_dev, args.steps, with_kernel_timing=True)
    launches = _lib.launch_count() - launches0
    k_ms, k_n = C.c_double(), C.c_int64()
    L.nerf_timing_read(0, C.byref(k_ms), C.byref(k_n))
    kernel_ms = {"mlp_fwd": k_ms.value / args.steps}
    for kind, nm in ((1, "mlp_bwd_chain"), (2, "wgrad")):
        a_ms, a_n = C.c_double(), C.c_int64()
        L.nerf_timing_read(kind, C.byref(a_ms), C.byref(a_n))
        if a_n.value:
            kernel_ms[nm] = a_ms.value / args.steps
    e2e_ms = timed(step_e2e, args.steps)
    clocks = sampler.stop()               # sampled under load over both timed regions (device-resident and end-to-end)

    extra = {}
    if args.mode == "train":
        # the metric names render AND train: time a few forward_pass steps on the same trainer / batches as well
        def step_render(i):
            img, o, d, t, u = dev_batches[i % R]
            return trainer.forward_pass(o, d, t, u_pdf=u, maps_only=True)[0][1]
        for i in range(3):
            step_render(i)
        r_ms = timed(step_render, max(3, args.steps // 2))
        L.nerf_timing_read(0, C.byref(C.c_double()), C.byref(C.c_int64()))
        extra["render_rays_per_sec"] = world * B / (r_ms / max(3, args.steps // 2) * 1e-3)
    ms_per_step = total_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)
    e2e_value = world * B / (e2e_ms / args.steps * 1e-3)
    h2d = sum(x.numel() * x.element_size() for x in host_batches[0])
    d2h = 12 if args.mode == "train" else B * 12

    # roofline of the dominant kernel: the fused tcgen05 MLP forward kernel (coarse + fine launches)
    peaks = measured_peaks()
    samples_per_step = B * (Nc + Nc + Nf)
    fwd_flop_per_step = samples_per_step * FLOP_PER_SAMPLE_FWD
    roofline = None
    if k_n.value > 0:
        avg_ms = k_ms.value / k_n.value
        flop_per_launch = fwd_flop_per_step * args.steps / k_n.value
        achieved = flop_per_launch / (avg_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "nerf_mlp_fwd_tc_kernel", "achieved": achieved,
                    "peak": peaks["tensor_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor_sustained"],
                    "traffic": NCU_TRAFFIC_PER_LAUNCH[args.mode], "traffic_unit": "bytes/launch (ncu, profiles/r1_*)",
                    "peak_source": peaks["source"] + " (sustained bf16)",
                    "avg_launch_ms": avg_ms, "launches": k_n.value,
                    "share_of_step": k_ms.value / total_ms}

    # the other two tensor kernels of a training step (same live CUDA-event timers), for the record:
    #   weight gradient: HBM-bound, reads every saved activation + dZ image once = 1 294 336 B per 128-sample tile
    #   dX chain: tensor-bound, 557 056 MAC per sample (dZ.W^T of ddir[:256], feature and layers 7..1; heads on CUDA cores)
    if args.mode == "train" and "wgrad" in kernel_ms:
        tiles = -(-B * Nc // 128) + -(-B * (Nc + Nf) // 128)
        wg_bytes = tiles * (655360 + 638976)
        extra["roofline_wgrad"] = {"bound": "hbm", "kernel": "nerf_wgrad_tc_kernel", "unit": "GB/s", "peak": peaks["hbm"],
                                   "achieved": wg_bytes / (kernel_ms["wgrad"] * 1e-3) / 1e9,
                                   "frac": wg_bytes / (kernel_ms["wgrad"] * 1e-3) / 1e9 / peaks["hbm"],
                                   "traffic": (8.589644e9 + 2.817732e9) / 2, "traffic_unit": "bytes/launch (ncu, profiles/r1_*)",
                                   "peak_source": peaks["source"] + " (HBM copy)",
                                   "avg_launch_ms": kernel_ms["wgrad"] / 2, "launches": 2 * args.steps,
                                   "share_of_step": kernel_ms["wgrad"] / ms_per_step}
        chain_flop = samples_per_step * 2 * (256 * 128 + 8 * 256 * 256)
        extra["roofline_chain"] = {"bound": "tensor", "kernel": "nerf_mlp_bwd_tc_kernel", "unit": "TFLOP/s",
                                   "peak": peaks["tensor_sustained"],
                                   "achieved": chain_flop / (kernel_ms["mlp_bwd_chain"] * 1e-3) / 1e12,
                                   "frac": chain_flop / (kernel_ms["mlp_bwd_chain"] * 1e-3) / 1e12 / peaks["tensor_sustained"],
                                   "share_of_step": kernel_ms["mlp_bwd_chain"] / ms_per_step}

    # `roofline` is the kernel with the largest share of the step; the others keep their own keys
    if roofline is not None and "roofline_wgrad" in extra:
        cands = {"roofline_fwd": roofline, "roofline_wgrad": extra["roofline_wgrad"], "roofline_chain": extra["roofline_chain"]}
        top = max(cands, key=lambda k: cands[k]["share_of_step"])
        extra.pop("roofline_wgrad"); extra.pop("roofline_chain")
        roofline = cands.pop(top)
        extra.update(cands)

    line = {"metric": metric, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config,
            "samples_per_sec": value * (Nc + Nc + Nf),
            "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "kernel_ms_per_step": kernel_ms}
    line.update(extra)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_arm(args.mode, conf, 3, 1, CPU_SAMPLE_RAYS)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
