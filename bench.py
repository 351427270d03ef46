"""bench.py — headline benchmark of the MichiGAN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload gen_fwd|train_step]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): generator-only forward of netG=spadeb (ngf 64), batch 8 per GPU,
512x512 synthetic inputs, train-mode statistics (the `generate_fake` of a training iteration:
SyncBN batch statistics incl. the cross-rank exchange, spectral-norm power iteration), no grad.
One "step" = one such forward over one batch.  Prints ONE JSON line (rank 0).

  value     images/s, whole job, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e       images/s through the public API (`Pix2PixModel(data, mode='inference')`-style call) with
            pinned HOST buffers: H2D of the data dict and D2H of the generated image inside the timed region
  roofline  the dominant kernel (fused SPADE gamma/beta implicit GEMM of up_3) timed alone, live
  cpu_baseline / --impl reference   the CPU oracle port of the reference path on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BATCH_PER_GPU = 8
SIZE = 512
G_FWD_GFLOP_PER_IMG = 1114.2  # SURVEY.md §8d / BASELINE.md §3 (2*MAC, convs only)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="gen_fwd", choices=["gen_fwd", "train_step"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return p, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm (oracle port)
def best_cpu_threads():
    """The CPU arm gets every host thread it can USE: a 3x3 conv of the generator's shape is timed at several intra-op
    thread counts (all cores down to 16) and the fastest is kept - on the 128-thread B200 hosts oversubscribing makes
    PyTorch's CPU conv 3x slower than 16..32 threads (profiles/r01_cpu_threads.log)."""
    import torch.nn.functional as F
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cands = sorted({max(1, ncpu), max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 32), min(ncpu, 16)})
    x = torch.randn(1, 128, 256, 256)
    w = torch.randn(128, 128, 3, 3)
    best, best_t = cands[-1], float("inf")
    for t in cands:
        torch.set_num_threads(t)
        F.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(2):
            F.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
    return best


def cpu_generator_forward_ips(steps, warmup, sample_images=1):
    """The reference's generator forward restated on the CPU (oracle/michigan_oracle.py), all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import michigan_oracle as orc
    from helpers import preprocessed, reference_layout_state
    cores = best_cpu_threads()
    torch.set_num_threads(cores)
    cfg = dict(ngf=64, ndf=64, size=SIZE, batch=sample_images, data_seed=1234)
    sd = reference_layout_state("G", cfg, 0)
    _, pre = preprocessed(cfg)
    opt = orc.default_opt(isTrain=True)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.generate_fake(sd, opt, pre, True, rng_k=25)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    return sample_images * len(times) / total, total / len(times) * 1e3, cores


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = max(1, min(a.steps, 5)), max(1, min(a.warmup, 1))
    ips, ms, cores = cpu_generator_forward_ips(steps, warm, 1)
    line = {
        "impl": "reference", "metric": "512x512 images/sec (generator forward)", "value": ips, "unit": "images/s",
        "n_gpus": a.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "generator-only forward (netG=spadeb ngf64, train-mode statistics, no grad), 512x512 synthetic",
                   "per_step_images": 1, "note": "CPU arm: bounded sample of 1 image per step on the host cores"},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d timed forwards of 1 image (batch 1) after %d warm-up" % (steps, warm)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ native arm
# DRAM traffic of one launch of the dominant kernel from the `ncu --set full` capture committed as
# profiles/r01_ncu_spade_gemm_f16_final.txt (dram__bytes_read.sum 817.4 MB + dram__bytes_write.sum 1028.9 MB at N=8);
# algorithmic bytes = actv fp16 268 MB + x fp32 268 MB + weights 0.6 MB read, bf16 hi+lo 537 MB written.
NCU_TRAFFIC_BYTES_N8 = 817.437696e6 + 1028.886e6


def dominant_kernel_roofline(batch):
    """Time the fused SPADE gamma/beta implicit GEMM of up_3.norm_0 alone, in the operand format the forward uses
    (precision mode mixed16: fp16 operands, fp32 accumulate, bf16 hi/lo output; mode tf32: TF32 operands):
    A = actv [N,512,512,128], N_gemm = 2*128, K = 9*128, SPADE epilogue reading x [N,256,256,128] (upsample folded)
    and writing h [N,512,512,128].  FLOPs per launch = 2 * N*512*512 * 1152 * 256 (SURVEY.md Appendix A rows
    G.up_3.norm_0.mlp_gamma + mlp_beta = 2 * 77.309 GFLOP per image)."""
    from michigan_b200 import ops, precision
    dev = "cuda"
    actv = torch.randn(batch, SIZE, SIZE, 128, device=dev)
    wg = torch.randn(128, 128, 3, 3, device=dev) / 34
    xs = torch.randn(batch, SIZE // 2, SIZE // 2, 128, device=dev)
    v = torch.ones(128, device=dev)
    flush = torch.empty(192 * 1024 * 1024 // 4, device=dev)
    if precision.mode() == "tf32":
        wp = ops.pack_weight_gb(wg, wg)
        kind = "kind::tf32"
        f = lambda: ops.conv_igemm(actv, wp, 128, 3, 3, 1, 1, act=2, round_out=True, spade=(xs, 1, v, v, v, v))
    else:
        a16 = actv.half()
        wp = ops.pack_weight_gb16(wg, wg)
        kind = "kind::f16"
        f = lambda: ops.conv_igemm(a16, wp, 128, 3, 3, 1, 1, act=2, a_fmt=ops.F16, spade=(xs, 1, v, v, v, v),
                                   out16=(ops.BF16, True), want_f32=False)
    for _ in range(3):
        f()
    times = []
    for _ in range(5):
        flush.zero_()  # evict L2 (126 MB) between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    flops = 2.0 * batch * SIZE * SIZE * 1152 * 256
    return flops / (ms * 1e-3) / 1e12, ms, flops, kind


def run_native(a):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from michigan_b200 import _lib
    from michigan_b200.options import make_opt
    from michigan_b200.pix2pix_model import Pix2PixModel
    from michigan_b200.synth import fill_state_dict, synthetic_batch

    if a.workload == "train_step":
        return run_train_step(a, rank, world, local)
    torch.manual_seed(0)
    opt = make_opt(is_train=True, gpu_ids=[local], batchSize=a.batch * world)
    model = Pix2PixModel(opt)
    fill_state_dict(model.netG.state_dict(), 0)  # random-init weights of the reference architecture (109.5 M params)
    model.netG.train()
    batch = a.batch
    data = synthetic_batch(batch, SIZE, 1234 + rank)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}
    host_out = torch.empty(batch, 3, SIZE, SIZE).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    with torch.no_grad():
        pre = model.preprocess_input(host)
        torch.cuda.synchronize()

        def step():
            return model.generate_fake(pre[0], pre[2], pre[4], pre[1], pre[3], pre[5])

        for _ in range(a.warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            out = step()
        e1.record()
        barrier()
        launches = _lib.launch_count() - l0
        clocks = sampler.stop() if rank == 0 else None
        ms_total = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        ms_total = ms_total.item()

        # ---- end to end through the public API with host buffers.  Every step copies its inputs from pinned host memory
        # and reads its result back into pinned host memory; like a prefetching data loader the copies run on a second
        # stream (double-buffered), so step i+1's H2D and step i-1's D2H overlap step i's kernels.
        copy_stream = torch.cuda.Stream()
        main_stream = torch.cuda.current_stream()
        host_outs = [host_out, torch.empty_like(host_out).pin_memory()]
        tensor_keys = [k for k, v in host.items() if torch.is_tensor(v)]

        # two device-side input sets allocated once (no allocator traffic on the copy stream)
        dev_sets = [{k: torch.empty_like(host[k], device="cuda") for k in tensor_keys} for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                for k in tensor_keys:
                    dev_sets[i & 1][k].copy_(host[k], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            dev = dict(host)
            dev.update(dev_sets[i & 1])
            return dev, ev

        def e2e_run(nsteps):
            nxt = prefetch(0)
            for i in range(nsteps):
                dev, ev = nxt
                main_stream.wait_event(ev)
                if i + 1 < nsteps:
                    nxt = prefetch(i + 1)   # queued behind step i-1's read-back, which waited for step i-1's kernels
                img = model(dev, mode="inference")
                done = torch.cuda.Event()
                done.record(main_stream)
                img.record_stream(copy_stream)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done)
                    host_outs[i & 1].copy_(img, non_blocking=True)
            copy_stream.synchronize()
            main_stream.synchronize()

        e2e_run(max(2, a.warmup // 2 + 1))
        barrier()
        e2e_steps = a.steps
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_e2e = t_e2e.item()

    h2d = sum(v.numel() * v.element_size() for k, v in host.items()
              if torch.is_tensor(v) and k in ("label_ref", "label_tag", "image_ref", "image_tag", "orient", "noise"))
    d2h = host_out.numel() * 4

    if rank == 0:
        pk, pk_kind = peaks()
        tf, kms, kflops, kkind = dominant_kernel_roofline(batch)
        from michigan_b200 import precision
        peak_tf = float(pk["bf16_tflops"])
        ms_step = ms_total / a.steps
        value = world * batch * a.steps / (ms_total * 1e-3)
        line = {
            "metric": "512x512 images/sec (generator forward)", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": DTYPE_NOTE[precision.mode()], "data": "synthetic",
            "config": {"workload": "generator-only forward (netG=spadeb ngf64, 109.5M params, train-mode statistics, no grad), "
                                   "batch %d/GPU, 512x512 synthetic mask/orient/ref inputs" % batch,
                       "global_batch": batch * world, "parallelism": "dp%d" % world,
                       "l2": "no explicit flush: each step streams multi-GB NHWC activations (>> 126 MB L2)",
                       "algorithmic_gflop_per_image": G_FWD_GFLOP_PER_IMG},
            "achieved_tflops_step": G_FWD_GFLOP_PER_IMG * batch / ms_step,
            "e2e": {"value": world * batch * e2e_steps / t_e2e, "unit": "images/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "how": "Pix2PixModel(data, mode='inference') per step; pinned host inputs -> device and image -> pinned host every "
                           "step, copies double-buffered on a second stream (prefetching-loader style), wall clock"},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"kernel": "igemm_tf32_kernel<1,16> (fused SPADE gamma|beta implicit GEMM + modulate + LeakyReLU, up_3.norm_0 shape, "
                                   "tcgen05 %s)" % kkind,
                         "bound": "tensor", "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                         "traffic": NCU_TRAFFIC_BYTES_N8 * batch / 8 if kkind == "kind::f16" else None,
                         "traffic_unit": "bytes per launch (dram read + write, ncu --set full, profiles/r01_ncu_spade_gemm_f16_final.txt)",
                         "peak_kind": "%s bf16 dense burst (MEASURED_PEAKS.json)%s" % (
                             pk_kind, "; kind::tf32 issues at half the bf16 rate" if kkind == "kind::tf32" else ""),
                         "ms_per_launch": kms, "flops_per_launch": kflops},
        }
        if not a.no_cpu_baseline and world == 1:
            ips, ms_cpu, cores = cpu_generator_forward_ips(3, 1, 1)
            line["cpu_baseline"] = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                    "sample": "3 timed forwards of 1 image (batch 1, same net/size) after 1 warm-up"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


DTYPE_NOTE = {
    "mixed16": "fp16/bf16 tensor-core operands (bf16 hi+lo split where needed), fp32 accumulate and storage",
    "tf32": "tf32",
}

TRAIN_GFLOP_PER_IMG = 4775.8  # G step + D step, hinge GAN + GAN-feature losses


def run_train_step(a, rank, world, local):
    """BASELINE.json configs[2]: full G+D train iteration (run_generator_one_step + run_discriminator_one_step),
    batch 8 per GPU, 512x512 synthetic; secondary metric (`--workload train_step`)."""
    import torch.distributed as dist
    from michigan_b200 import _lib
    from michigan_b200.options import make_opt
    from michigan_b200.synth import fill_state_dict, synthetic_batch
    from michigan_b200.trainer import Pix2PixTrainer
    torch.manual_seed(0)
    opt = make_opt(is_train=True, gpu_ids=[local], batchSize=a.batch * world, niter=50, niter_decay=0)
    trainer = Pix2PixTrainer(opt)
    m = trainer.pix2pix_model_on_one_gpu
    fill_state_dict(m.netG.state_dict(), 0)
    fill_state_dict(m.netD.state_dict(), 1)
    m.train()
    data = synthetic_batch(a.batch, SIZE, 1234 + rank)
    host = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in data.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        trainer.run_generator_one_step(dict(host))
        trainer.run_discriminator_one_step(dict(host))

    for _ in range(a.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / a.steps   # CPU time to issue one step (no sync inside)
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = ms_total.item()
    if rank == 0:
        ms_step = ms_total / a.steps
        losses = {k: float(v.mean()) for k, v in trainer.get_latest_losses().items()}
        h2d = sum(v.numel() * v.element_size() for k, v in host.items()
                  if torch.is_tensor(v) and k in ("label_ref", "label_tag", "image_ref", "image_tag", "orient", "noise")) * 2
        value = world * a.batch * a.steps / (ms_total * 1e-3)
        line = {
            "metric": "512x512 images/sec (train step)", "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "full G+D train iteration (hinge GAN + GAN-feature losses, Adam TTUR), netG=spadeb ngf64, "
                                   "netD=multiscale ndf64, batch %d/GPU, 512x512 synthetic" % a.batch,
                       "global_batch": a.batch * world, "parallelism": "dp%d" % world,
                       "l2": "no explicit flush: multi-GB activations per step", "algorithmic_gflop_per_image": TRAIN_GFLOP_PER_IMG},
            "achieved_tflops_step": TRAIN_GFLOP_PER_IMG * a.batch / ms_step,
            # the step already starts from host (pinned) tensors and returns host-visible loss scalars
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * len(losses)},
            "gpu_launches": launches, "clocks": clocks, "losses": losses, "host_enqueue_ms_per_step": host_enqueue_ms,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_native(args)
