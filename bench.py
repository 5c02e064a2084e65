"""Benchmark of the DIR inference hot path (BASELINE.json: images/sec at B=128, 256x256, 3 stages).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels through the C ABI)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's CPU forward

A step = one whole eval forward (ResNet-50 backbone -> init regression -> 2 refinement stages -> MANO
outputs, incl. the seg/dense/proj_feat heads the reference also computes) over one batch of B synthetic
images per GPU. For N>1 (torchrun, one rank per GPU) images shard over ranks (weak scaling) and each step
ends with the path's single collective, the all-gather of the output records.
Prints ONE JSON line on rank 0 (contract in the task statement). Keys beyond the contract:
  roofline          the dominant KERNEL = all conv_tc_kernel launches of the step (72 launches, ~78 % of the step):
                    executed algorithmic FLOPs / summed CUDA-event time, against the measured cuBLAS bf16 peak
  roofline_top      the largest single launch (InitRegressor attention conv), same method
  roofline_step     whole step: executed FLOPs (the 2560-channel fusion conv is factored away, so its 15.3 GFLOP/img
                    are NOT counted) and, labelled, the dense-equivalent figure of the reference's op sequence
  gpu_eager_baseline  the reference's op sequence in PyTorch eager on this GPU (fp32, TF32-allowed, bf16 autocast)
  parity            measured in this run on 8 of the benchmark's images against the CPU oracle
  no_aux            the same step without the seg/dense/proj_feat heads apps/eval.py never reads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_DENSE_EQUIV = 36.80  # GFLOP/img of the reference's own op sequence: incl. the 15.27 GF dense fusion convs
GF_JOINT_SPACE = 0.09   # grid_sample/pos-emb/SemGCN/mixSTE/regressor/MANO
GF_FUSION_FACTORED = 0.35  # what bone_coef + bone_fusion actually execute instead of the 15.27 GF dense convs
TOP_LAYER = "init_regressor.attention_left.0"
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
TF32_PEAK_TFLOPS = 891.0  # tcgen05 kind::tf32 issue rate measured by scripts/mma_probe.cu (profiles/mma_probe_r2.txt)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="images per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32", "tf32"])
    ap.add_argument("--cpu-batch", type=int, default=16, help="images per CPU-baseline forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip eager-GPU baseline, parity, no-aux and fp32-config legs")
    ap.add_argument("--cuda-graph", action="store_true")
    ap.add_argument("--backbone", default="resnet50", choices=["resnet50", "hrnet_w32", "hrnet_w48"],
                    help="resnet50 = the reference's network; hrnet_w32 / hrnet_w48 = extension, parity unpinned (BASELINE configs 3-5)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"burst": float(d.get("bf16_tflops", 1590.0)), "sustained": float(d.get("bf16_tflops_sustained", 1400.0)),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "measured (MEASURED_PEAKS.json)"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(s[2]) for s in self.samples)}


def cpu_forward_rate(batch, min_seconds, max_iters):
    """The reference's CPU implementation of the path: the oracle port (oracle/dir_oracle.py, pinned to the
    unmodified reference by tests/golden) with all host threads. Returns (images/s, cores, seconds, iters)."""
    import torch

    from oracle import dir_oracle as O
    from oracle.synth import make_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(0)
    img = torch.randn(batch, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    O.dir_forward(sd, img[: max(1, batch // 4)])  # warm-up
    t0 = time.perf_counter()
    iters = 0
    while iters < max_iters and (iters == 0 or time.perf_counter() - t0 < min_seconds):
        O.dir_forward(sd, img)
        iters += 1
    dt = time.perf_counter() - t0
    return batch * iters / dt, cores, dt, iters


def run_reference(args, rank):
    if rank != 0:
        return
    import torch

    from oracle import dir_oracle as O
    from oracle.synth import make_state_dict

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(0)
    bs = args.cpu_batch
    img = torch.randn(bs, 3, 256, 256, generator=torch.Generator().manual_seed(0))
    for _ in range(max(1, min(args.warmup, 2))):
        O.dir_forward(sd, img)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.dir_forward(sd, img)
    dt = time.perf_counter() - t0
    v = bs * args.steps / dt
    sample = (f"{args.steps} forwards of {bs} images (bounded sample of the B={args.batch} workload; CPU images/s is flat "
              f"above B~8, BASELINE.md 3), fp32, {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "images/sec", "value": v, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1000, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DIR eval forward, ResNet-50, 3 stages, 256x256, B={args.batch}/GPU (reference arm: "
                               f"CPU, {bs}-image sample per step)"},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_eager_baseline(dev, batch, steps=6):
    """BASELINE.md 3(4) / SURVEY 2: the reference's op sequence (oracle restatement: the same ATen / cuDNN / cuBLAS calls
    as models/dir.py, eager, no graphs) on this GPU. images/s for fp32 convs, TF32-allowed convs (torch's default) and
    bf16 autocast."""
    import torch

    from oracle import dir_oracle as O
    from oracle.synth import make_state_dict

    sd = {k: v.to(dev) for k, v in make_state_dict(0).items()}
    imgs = [torch.randn(batch, 3, 256, 256, device=dev) for _ in range(2)]
    out = {"batch": batch, "unit": "images/s", "how": "oracle/dir_oracle.py ops on CUDA, PyTorch eager, cudnn.benchmark"}
    prev = (torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32, ac in (("fp32", False, False), ("tf32_convs_torch_default", True, False), ("bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                for i in range(3):
                    O.dir_forward(sd, imgs[i % 2])
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    O.dir_forward(sd, imgs[i % 2])
                e1.record()
            torch.cuda.synchronize()
            out[name] = batch * steps / (e0.elapsed_time(e1) / 1000.0)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cudnn.benchmark = prev
    return out


def parity_check(net, precision, host_batch, n=8):
    """Parity of THIS run: n of the benchmark's own images through the net and through the CPU oracle."""
    import torch

    from oracle import dir_oracle as O
    from oracle.synth import make_state_dict

    sd = make_state_dict(0)
    img = host_batch[:n].clone()
    want = O.dir_forward(sd, img)
    outs, _ = net({"img": img}, None, None)
    torch.cuda.synchronize()

    def drift(a, b, i):
        d = torch.cat([(a[i][k].float().cpu() - b[i][k].float()).norm(dim=-1).flatten()
                       for k in ("pd_mesh_xyz_left", "pd_mesh_xyz_right")]) * 1000
        return float(d.mean()), float(d.max())

    worst = max(float((outs[i][k].float().cpu() - want[i][k]).abs().max() / (want[i][k].abs().max() + 1e-30))
                for i in range(3) for k in O.OUT_KEYS)
    res = {"against": f"CPU oracle (fp32 restatement of the reference, pinned by tests/golden) on {n} of the timed images",
           "worst_relative_error": worst, "stage2_mesh_drift_mm_mean": drift(outs, want, 2)[0],
           "stage2_mesh_drift_mm_max": drift(outs, want, 2)[1]}
    if precision == "bf16":
        with torch.autocast("cpu", dtype=torch.bfloat16):
            auto = O.dir_forward(sd, img)
        auto = [{k: (v.float() if v is not None else None) for k, v in d.items()} for d in auto[:3]]
        res["reference_bf16_autocast_stage2_mesh_drift_mm_mean"] = drift(auto, want, 2)[0]
        res["note"] = ("bf16 feature maps: the yardstick is the reference's own drift under torch.autocast(bfloat16) on the "
                       "same images; precision='fp32' meets 1e-4 (parity_fp32)")
    return res


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    import dir_b200
    from dir_b200 import capi
    from dir_b200.dist import bind_to_gpu_numa_node
    from oracle.synth import make_state_dict

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # pinned staging buffers must live on the GPU's own NUMA node: bind BEFORE anything is allocated (first touch)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)

    def make_net(precision, aux, max_batch=B, backbone=None):
        backbone = backbone or args.backbone
        n = dir_b200.DIR(21, "./misc/mano", precision=precision, aux_outputs=aux, max_batch=max_batch,
                         use_cuda_graph=args.cuda_graph, backbone=backbone).to(dev)
        n.load_state_dict(make_state_dict(0, backbone=backbone), strict=False)
        n.eval()
        return n

    net = make_net(args.precision, True)
    if world > 1:
        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        net.init_nccl(rank, world, bcast)

    NBUF = 4  # 4 x 100 MB fp32 batches > 126 MB L2: the input is never L2-resident between steps
    gen = torch.Generator().manual_seed(1000 + rank)
    host = [torch.randn(B, 3, 256, 256, generator=gen).pin_memory() for _ in range(NBUF)]
    resident = [h.to(dev) for h in host]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warm):
        for i in range(warm):
            fn(i)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_of(model):
        def step(i):
            o = model.run_raw(resident[i % NBUF])
            if world > 1:
                return model.allgather_records(o["record"])
            return o["record"]
        return step

    # ---- timed region: K steps, inputs resident in HBM, device timers, max over ranks
    step = step_of(net)
    for i in range(W):
        step(i)
    sync_all()
    h = net._handle
    launches_per_step = h.lib.dirb200_forward_launches(h.h, B) + (1 if world > 1 else 0)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms = timed(step, K, 0)
    if sampler:
        sampler.stop_flag.set()
        sampler.join()
    value = world * B * K / (ms / 1000.0)

    # ---- per-conv CUDA events (launch stream) over a second, shorter run of the same step: the dominant kernel's time
    conv_rows, prof_steps = [], 0
    if not args.cuda_graph and rank == 0:
        prof_steps = max(2, min(K, 10))
        h.profile_layer("")
        for i in range(prof_steps):
            net.run_raw(resident[i % NBUF])  # the forward alone: rank 0 must not enter a collective on its own
        torch.cuda.synchronize()
        conv_rows = h.profile_dump()
        h.profile_read()
        h.profile_layer(None)
    if world > 1:
        dist.barrier()

    # ---- end to end through the public API (DIR.forward): pinned HOST input -> H2D -> forward -> D2H of the record
    e2e = e2e_f32 = None
    h2d_gbs = None
    if not args.no_e2e:
        out_hosts = [torch.empty(B, capi.RECORD_FLOATS).pin_memory() for _ in range(2)]
        d2h = torch.cuda.Stream(device=dev)  # the caller's download stream: result i leaves while forward i+1 runs
        Ke = max(3, K)
        frames = [torch.randint(0, 256, (B, 256, 256, 3), dtype=torch.uint8, generator=gen).pin_memory()
                  for _ in range(NBUF)]

        def download(i, rec):
            ev = torch.cuda.Event()
            ev.record()
            d2h.wait_event(ev)
            with torch.cuda.stream(d2h):
                out_hosts[i % 2].copy_(rec, non_blocking=True)
            rec.record_stream(d2h)

        def e2e_step_of(inputs):
            def f(i):
                outs, _ = net({"img": inputs[i % NBUF]}, None, None)  # forward() does the H2D copy (models/dir.py:514)
                rec = outs[0]["pd_mesh_xyz_left"]._base  # the stage dicts are views into one packed record
                if world > 1:
                    rec = net.allgather_records(rec)[rank * B:(rank + 1) * B]
                download(i, rec)
            return f

        def run_e2e(inputs, bytes_in, label):
            f = e2e_step_of(inputs)
            for i in range(3):
                f(i)
            sync_all()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h0 = time.perf_counter()
            host_ms = 0.0
            for i in range(Ke):
                f(i)
                if i == 4:  # 5 steps = ~550 launches: the launch queue cannot be full yet, so this is pure host work
                    host_ms = (time.perf_counter() - h0) * 1000.0 / 5
            torch.cuda.current_stream().wait_stream(d2h)  # the last result must have reached the host inside the region
            e1.record()
            sync_all()
            t = torch.tensor([e0.elapsed_time(e1), host_ms], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return {"value": world * B * Ke / (float(t[0].item()) / 1000.0), "unit": "images/s",
                    "h2d_bytes_per_step": bytes_in, "d2h_bytes_per_step": B * capi.RECORD_FLOATS * 4, "steps": Ke,
                    "input": label, "ms_per_step": float(t[0].item()) / Ke,
                    "host_enqueue_ms_per_step_max_over_ranks": float(t[1].item())}

        # the headline e2e: what a camera / cv2 hands over (apps/eval.py:56-61 runs the normalisation on the HOST per
        # sample; here it runs on the device, fused into the stem operand packing) — 4x fewer H2D bytes than fp32
        e2e = run_e2e(frames, B * 256 * 256 * 3, "uint8 HWC BGR host frames (pinned); BGR->RGB,/255,mean/std on the device")
        e2e_f32 = run_e2e(host, B * 3 * 256 * 256 * 4, "fp32 NCHW normalised host images (pinned): the tensor "
                                                       "apps/eval.py:168 passes")
        # H2D bandwidth this rank sees while all ranks copy at once (names the limiter of e2e scaling)
        sync_all()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(8):
            resident[i % NBUF].copy_(host[i % NBUF], non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        t = torch.tensor([c0.elapsed_time(c1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d_gbs = 8 * B * 3 * 256 * 256 * 4 / (float(t.item()) / 1000.0) / 1e9

    extras = world == 1 and not args.no_extras
    no_aux = parity = parity32 = fp32_cfg = eager = hrnet_cfg = hrnet48_cfg = None
    if extras and args.backbone != "resnet50":
        extras = False  # the extra legs (oracle parity, eager baseline) are defined for the reference's network
    if extras:
        net_na = make_net(args.precision, False)
        ms_na = timed(step_of(net_na), K, W)
        no_aux = {"value": B * K / (ms_na / 1000.0), "unit": "images/s", "ms_per_step": ms_na / K,
                  "what": "aux_outputs=False: without conv_final/seg/dense/proj_feat, which apps/eval.py:170-172 never reads"}
        del net_na
        parity = parity_check(net, args.precision, host[0])
        if args.precision == "bf16":  # the parity configuration (BASELINE configs[1]) beside the headline, same run
            net32 = make_net("fp32", True, 32)
            res32 = [r[:32].contiguous() for r in resident]
            ms32 = timed(lambda i: net32.run_raw(res32[i % NBUF]), max(5, K // 2), W)
            fp32_cfg = {"value": 32 * max(5, K // 2) / (ms32 / 1000.0), "unit": "images/s", "batch": 32,
                        "ms_per_step": ms32 / max(5, K // 2),
                        "what": "precision='fp32' (3xTF32 tcgen05 convs, round-to-nearest accumulation), B=32: BASELINE configs[1]"}
            parity32 = parity_check(net32, "fp32", host[0])
            del net32
        eager = gpu_eager_baseline(dev, B if B <= 128 else 128)
        # BASELINE.json configs[2] names HRNet-W32 as the headline network; the reference has none. The
        # extension (self-authored oracle, parity UNPINNED) is timed here so that the config has a number, labelled as such
        hrnet_legs = {}
        for bb in ("hrnet_w32", "hrnet_w48"):
            net_h = make_net(args.precision, True, B, bb)
            ms_h = timed(step_of(net_h), max(5, K // 2), W)
            hh = net_h._handle
            hrnet_legs[bb] = {"value": B * max(5, K // 2) / (ms_h / 1000.0), "unit": "images/s", "batch": B,
                              "ms_per_step": ms_h / max(5, K // 2),
                              "gpu_launches_per_step": hh.lib.dirb200_forward_launches(hh.h, B),
                              "what": f"backbone='{bb}', precision='{args.precision}': EXTENSION, not in the reference; parity "
                                      "unpinned (tests/test_gpu_hrnet.py checks it against the self-authored "
                                      "oracle/hrnet_oracle.py)"}
            del net_h
        hrnet_cfg, hrnet48_cfg = hrnet_legs["hrnet_w32"], hrnet_legs["hrnet_w48"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    mma_per_mac = {"bf16": 1, "tf32": 1, "fp32": 3}[args.precision]
    tensor_peak = pk["burst"] if args.precision == "bf16" else TF32_PEAK_TFLOPS / mma_per_mac
    peak_note = ("cuBLAS bf16 burst, MEASURED_PEAKS.json" if args.precision == "bf16" else
                 f"tcgen05 kind::tf32 issue rate 891 TFLOP/s (profiles/mma_probe_r2.txt) / {mma_per_mac} MMAs per MAC")
    roof = roof_top = None
    conv_gf_img = None
    if conv_rows:
        tc_rows = [r for r in conv_rows if r["tc"]]
        fam_ms = sum(r["ms"] for r in tc_rows)
        fam_fl = sum(r["flops"] for r in tc_rows)
        n_launch = len(tc_rows)
        conv_gf_img = sum(r["flops"] for r in conv_rows) / prof_steps / B / 1e9
        ach = fam_fl / (fam_ms / 1000.0) / 1e12
        traffic = None
        if os.path.exists(TRAFFIC_FILE):  # dram bytes per launch (family average) from the committed ncu capture
            with open(TRAFFIC_FILE) as f:
                t = json.load(f)
            if t.get("batch") == B and t.get("precision") == args.precision:
                traffic = t.get("family_dram_bytes_per_launch")
        kname = "conv_tc_kernel<BN,128,CG,PRE> + conv3x3_c64_halo_kernel + conv1x1_b2b_kernel<N2> (tcgen05 bf16 implicit GEMM)" if args.precision == "bf16" else \
                "conv_tf32_kernel<BN> (tcgen05 kind::tf32 implicit GEMM)"
        roof = {"bound": "tensor", "kernel": f"{kname}: all {n_launch // prof_steps} launches of a step "
                                             f"({fam_fl / n_launch / 1e9:.1f} GFLOP/launch executed, algorithmic 2*M*N*K)",
                "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak, "traffic": traffic,
                "peak_source": f"{peak_note}; {pk['source']}",
                "frac_of_sustained_peak": ach / pk["sustained"] if args.precision == "bf16" else None,
                "algorithmic_bytes_per_launch": sum(r["bytes"] for r in tc_rows) / n_launch,
                "launches_timed": n_launch, "avg_launch_ms": fam_ms / n_launch,
                "share_of_step": fam_ms / prof_steps / (ms / K),
                "how": f"CUDA events on the launch stream around every launch, {prof_steps} steps after the timed region"}
        top = [r for r in tc_rows if r["layer"].startswith(TOP_LAYER)]
        if top:
            t_ms, t_fl = sum(r["ms"] for r in top), sum(r["flops"] for r in top)
            a2 = t_fl / (t_ms / 1000.0) / 1e12
            roof_top = {"bound": "tensor", "kernel": f"largest single launch: {TOP_LAYER} (conv3x3 2048->2x1024 @8x8, "
                                                     f"{t_fl / len(top) / 1e9:.1f} GFLOP)",
                        "achieved": a2, "peak": tensor_peak, "unit": "TFLOP/s", "frac": a2 / tensor_peak,
                        "avg_launch_ms": t_ms / len(top), "share_of_step": t_ms / prof_steps / (ms / K)}
    step_roof = None
    if world == 1 and conv_gf_img is not None:
        gf_exec = conv_gf_img + GF_FUSION_FACTORED + GF_JOINT_SPACE
        tf = gf_exec * B * K / 1000.0 / (ms / 1000.0)
        step_roof = {"bound": "tensor", "achieved": tf, "peak": tensor_peak, "unit": "TFLOP/s", "frac": tf / tensor_peak,
                     "gflop_per_image_executed": gf_exec,
                     "note": f"whole step, EXECUTED FLOPs: convs {conv_gf_img:.2f} (live profile) + factored fusion "
                             f"{GF_FUSION_FACTORED} + joint space {GF_JOINT_SPACE} GFLOP/img; the 15.27 GFLOP/img of the dense "
                             f"2560-channel fusion convs are factored away and not credited",
                     "dense_equivalent": {"gflop_per_image": GF_DENSE_EQUIV,
                                          "tflops": GF_DENSE_EQUIV * B * K / 1000.0 / (ms / 1000.0),
                                          "label": "reference op sequence incl. work this implementation never executes"}}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, dt, iters = cpu_forward_rate(args.cpu_batch, 10.0, 12)
        cpu = {"value": v, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{iters} forwards of {args.cpu_batch} images in {dt:.1f} s (same weights/input recipe; CPU "
                         f"images/s is flat above B~8), fp32"}
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16", "fp32": "f32", "tf32": "tf32"}[args.precision], "data": "synthetic",
        "config": {"workload": f"DIR eval forward (ResNet-50 backbone, init regression, 2 refinement stages, "
                               f"seg/dense/proj_feat heads), 256x256, B={B} per GPU, random-init weights + synthetic MANO"
                               + ("" if args.backbone == "resnet50" else f" [--backbone {args.backbone}: extension, ResNet-50 replaced]"),
                   "global_batch": B * world, "parallelism": f"dp{world}" if world > 1 else "single",
                   "l2": f"{NBUF} rotating resident input batches (4x100 MB > 126 MB L2); activations ~GBs per step",
                   "collective": "ncclAllGather of (B,14661) fp32 records per step" if world > 1 else None,
                   "numa": numa},
        "clocks": sampler.summary() if sampler else None,
        "e2e": e2e, "e2e_fp32_input": e2e_f32, "h2d_gbs_per_rank_all_ranks_copying": h2d_gbs,
        "gpu_launches": launches_per_step * K,
        "roofline": roof, "roofline_top": roof_top, "roofline_step": step_roof,
        "cpu_baseline": cpu, "gpu_eager_baseline": eager, "parity": parity, "no_aux": no_aux,
        "fp32_config": fp32_cfg, "parity_fp32": parity32, "hrnet_w32_extension": hrnet_cfg, "hrnet_w48_extension": hrnet48_cfg,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
