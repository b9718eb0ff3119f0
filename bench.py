#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json): QK-Skip self-attention forward at the
Wan2.1-14B shape (B=1 per GPU, S=75600, H=40, D=128, bf16) at 42 % tile sparsity, one step = one pass of the
hot path (skip-list-gated forward + skip-list update) over one batch of synthetic Q/K/V.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sparsity 0.42]
    torchrun ... bench.py --gpus N ...          (one rank per GPU, batch-parallel, O gathered to rank 0 over NCCL)

Prints ONE JSON line (rank 0).  `value` = effective TFLOP/s = dense FLOPs (4*B*H*S^2*D, summed over ranks) /
step time, inputs resident in HBM.  `e2e` = the same through LiteAttention.__call__ with pinned HOST buffers
(H2D of q/k/v and D2H of O inside the timed region, double-buffered).  `roofline` = executed tensor FLOPs of
la_fwd_kernel per launch / its CUDA-event duration against the measured bf16 peak.  `cpu_baseline` / `--impl
reference` = torch SDPA on the host cores on a bounded slice of the same workload (the reference's CPU path
has no sparse mode: it always does the dense problem).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S_WAN, H_WAN, D_WAN = 75600, 40, 128          # 21 x 45 x 80 tokens (720p x 81 frames), Wan2.1/2.2-14B
METRIC = "self_attn_effective_tflops_s75600_d128_sparsity42"
UNIT = "TFLOP/s (dense-equivalent: 4*B*H*S^2*D / step time)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sparsity", type=float, default=0.42)
    ap.add_argument("--seq", type=int, default=S_WAN)
    ap.add_argument("--heads", type=int, default=H_WAN)
    ap.add_argument("--groups", type=int, default=0, help="head groups for the gather pipeline (0 = auto)")
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: peer = forward epilogue stores O into rank 0's symmetric buffer over NVLink (fused); "
                         "nccl = NCCL gather pipelined by head group")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-comparators", action="store_true", help="skip the dense GPU comparators (N=1 only)")
    ap.add_argument("--no-traffic", action="store_true", help="skip the live ncu DRAM-traffic measurement (N=1 only)")
    ap.add_argument("--no-seqpar", action="store_true", help="N>1: skip the sequence-parallel (one prompt over N GPUs) leg")
    ap.add_argument("--sweep", action="store_true", help="time sparsity 0/21/42/57/77 % instead of 0/42/77 % (kernel only)")
    ap.add_argument("--must-do", action="store_true", help="also time the update kernel with a must-do list (text tokens)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sdpa_sample(seq, steps, warmup, rows=4096):
    """The reference's CPU path for this op: torch.nn.functional.scaled_dot_product_attention on the host cores.
    Bounded sample: 1 head, `rows` query rows against all `seq` keys, d=128, fp32 math (dense; no sparse mode)."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    q = torch.randn(1, 1, rows, D_WAN, generator=g)
    k = torch.randn(1, 1, seq, D_WAN, generator=g)
    v = torch.randn(1, 1, seq, D_WAN, generator=g)
    for _ in range(warmup):
        F.scaled_dot_product_attention(q, k, v)
    t0 = time.perf_counter()
    for _ in range(steps):
        F.scaled_dot_product_attention(q, k, v)
    dt = (time.perf_counter() - t0) / steps
    flops = 4.0 * rows * seq * D_WAN
    return {"tflops": flops / dt / 1e12, "sec_per_sample": dt, "cores": torch.get_num_threads(),
            "sample": f"torch SDPA (CPU, fp32, dense): 1 head x {rows} query rows x {seq} keys x d=128 per step, "
                      f"{steps} steps after {warmup} warm-up"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)      # a step = the bounded slice below (~70 ms on 16+ cores)
    r = cpu_sdpa_sample(args.seq, steps, warmup)
    dense = 4.0 * args.heads * args.seq * float(args.seq) * D_WAN
    line = {
        "impl": "reference", "metric": METRIC, "value": r["tflops"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup,
        "ms_per_step": dense / (r["tflops"] * 1e12) * 1e3,      # extrapolated to one full B=1 call
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Wan2.1-14B self-attention shape B=1 S={args.seq} H={args.heads} D=128, dense on CPU "
                               "(the reference's CPU path, torch SDPA, has no QK-Skip mode); bounded slice, see "
                               "cpu_baseline.sample"},
        "cpu_baseline": {"value": r["tflops"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                         "sample": r["sample"]},
        "e2e": {"value": r["tflops"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-i", str(self.gpu), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001
            self.p = None

    def stop(self):
        if self.p is None:
            return {"error": "nvidia-smi unavailable"}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
                pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"error": "no samples"}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(gpu_index):
    """Best effort: run this process (and so first-touch its pinned buffers) on the CPUs of the GPU's NUMA node."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(gpu_index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus              # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "unknown (numa_node = -1)"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.extend(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & set(cpus)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{node} (bound to {len(allowed)} cpus)"
        return f"{node} (not bound: no allowed cpu on that node)"
    except Exception as e:  # noqa: BLE001
        return f"unknown ({type(e).__name__})"


def measure_dram_traffic(args):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE la_fwd_kernel launch at this run's configuration, measured by
    running tools/one_launch.py under ncu in a subprocess (after the timed regions; never timed).  Returns (bytes, source)
    or (None, reason)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    script = os.path.join(ROOT, "tools", "one_launch.py")
    if not os.path.exists(ncu) or not os.path.exists(script):
        return None, "ncu or tools/one_launch.py not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k",
           "regex:la_fwd_kernel", "-s", "2", "-c", "1", "--csv", sys.executable, script, "--seq", str(args.seq),
           "--heads", str(args.heads), "--sparsity", str(args.sparsity)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300).stdout
        tot = 0.0
        found = 0
        for ln in out.splitlines():
            if "dram__bytes_" in ln:
                cols = [c.strip().strip('"') for c in ln.split('","')]
                unit, val = cols[-2].lower(), float(cols[-1].replace(",", ""))
                tot += val * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
                found += 1
        if found >= 2:
            return tot, "ncu subprocess in this run (one launch, dram__bytes_read.sum + dram__bytes_write.sum)"
        return None, "ncu produced no dram__bytes rows: " + out[-160:].replace("\n", " | ")
    except Exception as e:  # noqa: BLE001
        return None, f"{type(e).__name__}: {str(e)[:120]}"


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from liteattention_b200 import LiteAttention, _native, synth
    from liteattention_b200.lite_attention import host_head_groups
    from liteattention_b200.dist import BatchParallelLiteAttention

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    B, S, H, D = 1, args.seq, args.heads, D_WAN
    qt, kt = synth.tile_counts(S)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    q, k, v = (torch.randn(B, S, H, D, device=dev, generator=g).to(torch.bfloat16) for _ in range(3))
    dense_flops = synth.flops_dense(B, H, S, S, D)

    def make_list(sp, heads, seed):
        if sp <= 0:
            return LiteAttention.init_skip_list(B, S, heads, D, False, torch.bfloat16, dev)[0]
        rl, _ = synth.exact_sparsity_list(B, heads, qt, kt, sp, seed=seed, device=dev)
        return rl

    # ---- kernel-level timing: forward + update as two C-ABI calls with CUDA events around each ------------
    def time_kernels(sp, steps, warmup, rl=None, must_do=None):
        rl = make_list(sp, H, seed=1234) if rl is None else rl
        wl = torch.zeros_like(rl)
        out = torch.empty_like(q)
        lse = torch.empty(B, H, S, device=dev, dtype=torch.float32)
        stat = torch.empty(B, H, qt, kt, device=dev, dtype=torch.float32)
        scale = D ** -0.5
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        for _ in range(warmup):
            _native.fwd(q, k, v, out, lse, scale, rl, stat)
            _native.skip_update(rl, must_do, wl, stat, B, H, qt, kt, float("-inf"))
        torch.cuda.synchronize()
        for i in range(steps):
            ev[i][0].record()
            _native.fwd(q, k, v, out, lse, scale, rl, stat)
            ev[i][1].record()
            _native.skip_update(rl, must_do, wl, stat, B, H, qt, kt, float("-inf"))
            ev[i][2].record()
        torch.cuda.synchronize()
        assert torch.equal(wl[..., 0], rl[..., 0])              # thr = -inf: the list is stationary
        fwd_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in ev)
        upd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
        return {"sparsity": LiteAttention.sparsity(rl), "fwd_ms": fwd_ms, "update_ms": upd_ms,
                "exec_flops": synth.flops_executed(rl, S, S, D), "update_bytes": synth.update_bytes(rl, wl), "rl": rl}

    # ---- the step the contract times: public objects, batch-parallel, O gathered to rank 0 ----------------
    n_groups = args.groups if args.groups > 0 else (1 if world == 1 else 5)

    def factory():
        return LiteAttention(enable_skipping=True, threshold=float("-inf"), max_batch_size=B)
    bp = BatchParallelLiteAttention(factory, num_heads=H, num_groups=n_groups, dst=0,
                                    peer_store=(args.gather == "peer" and world > 1))
    for gi, gsl in enumerate(bp.groups):                         # preset every head group's list at the target sparsity
        bp.attn[gi].load_skip_list(make_list(args.sparsity, gsl.stop - gsl.start, seed=1234 + gi), q[:, :, gsl], v[:, :, gsl])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        bp(q, k, v)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        bp(q, k, v)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _native.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    stats_t = torch.tensor([ms, float(launches)], device=dev, dtype=torch.float64)
    if world > 1:
        mx = stats_t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats_t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, launches = float(mx[0]), int(sm[1])
    step_sparsity = statistics.mean(a.last_sparsity(B) for a in bp.attn)

    # ---- N>1: did rank 0 receive what every rank computed?  (driver-visible correctness of the gather) ----
    gather_verified = None
    if world > 1:
        def checksum(t):                                        # order-sensitive 2 x int64 digest of a bf16 tensor
            x = t.contiguous().view(torch.int16).to(torch.int64).view(-1)
            w = (torch.arange(x.numel(), device=x.device, dtype=torch.int64) % 8191) + 1
            return torch.stack([x.sum(), (x * w).sum()])
        _, gathered = bp(q, k, v)                               # thr = -inf: lists are stationary, O repeats bit for bit
        torch.cuda.synchronize()
        local = torch.empty_like(q)                             # the same call into a LOCAL buffer, one head group at a time
        for gi, gsl in enumerate(bp.groups):
            chk = factory()
            chk.load_skip_list(bp.attn[gi].read_list[:B].clone(), q[:, :, gsl], v[:, :, gsl])
            local[:, :, gsl] = chk(q[:, :, gsl], k[:, :, gsl], v[:, :, gsl])
        mine = checksum(local)
        allsums = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allsums, mine)
        if rank == 0:
            ok = True
            for r in range(world):
                slab = torch.cat([gathered[gi][r] for gi in range(len(gathered))], dim=2) if len(gathered) > 1 else gathered[0][r]
                ok &= bool(torch.equal(checksum(slab), allsums[r]))
            gather_verified = ok
        del local
        barrier()

    # ---- N>1: one prompt over N GPUs (sequence-parallel): all_to_all in, O scattered by the epilogue ----
    seqpar = None
    if world > 1 and not args.no_seqpar and S % world == 0 and H % world == 0:
        try:
            from liteattention_b200.dist import UlyssesLiteAttention
            sl_ = S // world
            gsp = torch.Generator(device=dev).manual_seed(4242)              # same seed on every rank: one shared prompt
            qf, kf, vf = (torch.randn(B, S, H, D, device=dev, generator=gsp).to(torch.bfloat16) for _ in range(3))
            loc = slice(rank * sl_, (rank + 1) * sl_)
            ql, kl, vl = (t[:, loc].contiguous() for t in (qf, kf, vf))
            ul = UlyssesLiteAttention(lambda: LiteAttention(enable_skipping=False, max_batch_size=B))
            for _ in range(2):
                o_sp = ul(ql, kl, vl)
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            nsp = max(3, min(args.steps, 10))
            for _ in range(nsp):
                o_sp = ul(ql, kl, vl)
            s1.record()
            barrier()
            sp_ms = torch.tensor([s0.elapsed_time(s1) / nsp], device=dev, dtype=torch.float64)
            dist.all_reduce(sp_ms, op=dist.ReduceOp.MAX)
            # correctness: my tokens, the heads of rank (rank+1)%world, against a single-GPU dense call on those heads
            hl = H // world
            hsl = slice(((rank + 1) % world) * hl, ((rank + 1) % world + 1) * hl)
            ref_o = LiteAttention(enable_skipping=False, max_batch_size=B)(qf[:, :, hsl], kf[:, :, hsl], vf[:, :, hsl])
            okt = torch.tensor([int(torch.equal(o_sp[:, :, hsl], ref_o[:, loc]))], device=dev)   # same Q tiling => bit-exact
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
            seqpar = {"ms_per_call": float(sp_ms[0]), "verified_bit_exact_vs_single_gpu": bool(int(okt[0])),
                      "config": f"one dense Wan-shape prompt (S={S}, H={H}) sequence-sharded over {world} GPUs: all_to_all of "
                                "q/k/v (NCCL), forward on H/N heads, O rows stored by the epilogue into the owning rank's "
                                "symmetric buffer (peer stores over NVLink)",
                      "effective_tflops": dense_flops / (float(sp_ms[0]) * 1e-3) / 1e12}
            del qf, kf, vf, ql, kl, vl, ul, o_sp
        except Exception as e:  # noqa: BLE001 -- auxiliary leg; the headline does not depend on it
            seqpar = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        barrier()

    # ---- end to end through LiteAttention.__call__ with pinned host buffers ---------------------------------
    e2e = None
    if not args.no_e2e:
        numa_note = bind_to_gpu_numa_node(local_rank)            # pinned buffers land on the GPU's own NUMA node
        hq, hk, hv = (t.cpu().pin_memory() for t in (q, k, v))
        la = factory()
        la.load_skip_list(make_list(args.sparsity, H, seed=1234), q, v)
        main_s = torch.cuda.current_stream()
        # The call a user with host-resident (offloaded) activations makes: LiteAttention.__call__ on pinned CPU tensors
        # returns O in pinned host memory; inside, q/k/v go up and O comes down by head groups around the per-group
        # forward + list update (liteattention_b200/lite_attention.py:_call_host).  Every step uploads its own inputs
        # and downloads its own result inside the timed region.
        ho = None
        for i in range(max(2, min(args.warmup, 3))):
            ho = la(hq, hk, hv)
        la.join_host_copies()
        torch.cuda.synchronize()
        barrier()
        e0.record()
        for i in range(args.steps):
            ho = la(hq, hk, hv)
        la.join_host_copies()
        e1.record()
        barrier()
        dbuf = [[torch.empty_like(q) for _ in range(3)]]
        ho = ho if ho is not None else torch.empty(B, S, H, D, dtype=torch.bfloat16).pin_memory()
        e2e_ms = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t[0])
        # what the copies alone achieve on this host with all ranks copying at once (the e2e ceiling when N ranks share it)
        barrier()
        e0.record()
        for _ in range(3):
            for dst_, src_ in zip(dbuf[0], (hq, hk, hv)):
                dst_.copy_(src_, non_blocking=True)
        e1.record()
        barrier()
        h2d_gbs = 3 * 3 * q.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        e0.record()
        for _ in range(3):
            ho.copy_(dbuf[0][0], non_blocking=True)
        e1.record()
        barrier()
        d2h_gbs = 3 * q.numel() * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
        bw = torch.tensor([h2d_gbs, d2h_gbs], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(bw, op=dist.ReduceOp.SUM)
        copy_floor_ms = max(3 * q.numel() * 2 * world / (float(bw[0]) * 1e9), q.numel() * 2 * world / (float(bw[1]) * 1e9)) * 1e3
        e2e = {"value": dense_flops * world / (e2e_ms * 1e-3) / 1e12, "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": 3 * q.numel() * 2 * world, "d2h_bytes_per_step": q.numel() * 2 * world,
               "api": "LiteAttention.__call__(pinned host q, k, v) -> pinned host O; uploads / forward + list update / downloads pipelined by head groups inside the call (" + str(host_head_groups(H)) + " from idle, " + str(host_head_groups(H, True)) + " while the previous call is in flight), two staging slots across calls",
               "host_link": {"aggregate_h2d_gbs": float(bw[0]), "aggregate_d2h_gbs": float(bw[1]),
                             "copy_only_floor_ms_per_step": copy_floor_ms, "numa_node": numa_note,
                             "limiter": ("host<->device copies (PCIe / host memory), all ranks through one host"
                                         if copy_floor_ms > 0.8 * ms else "kernel")}}
        del hq, hk, hv, ho, dbuf

    # ---- roofline of the dominant kernel (rank 0), sweep, CPU baseline -------------------------------------
    line = None
    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak_sust = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_burst = peaks.get("bf16_tflops", 1590.0)
        kt_main = time_kernels(args.sparsity, args.steps, args.warmup)
        achieved = kt_main["exec_flops"] / (kt_main["fwd_ms"] * 1e-3) / 1e12
        traffic, traffic_src = None, None
        if world == 1 and not args.no_traffic:
            traffic, traffic_src = measure_dram_traffic(args)       # one launch under ncu, after the timed regions
        if traffic is None:
            tr_path = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tr_path):
                tj = json.load(open(tr_path)).get("la_fwd_kernel", {})
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "profiles/traffic.json (" + str(tj.get("source", "earlier ncu capture")) + ")" + \
                              ("" if traffic_src is None else "; live measurement failed: " + traffic_src)
        roofline = {"bound": "tensor", "kernel": "la_fwd_kernel", "achieved": achieved, "peak": peak_sust,
                    "unit": "TFLOP/s", "frac": achieved / peak_sust, "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": 4.0 * q.numel() * 2 + B * H * S * 4 + 2.0 * B * H * qt * (kt + 1) * 4,
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
                                    if peaks else "fallback"),
                    "frac_of_burst_peak": achieved / peak_burst, "kernel_ms": kt_main["fwd_ms"],
                    "exec_flops_per_launch": kt_main["exec_flops"],
                    "update_kernel": {"ms": kt_main["update_ms"], "bound": "hbm",
                                      "achieved": kt_main["update_bytes"] / (kt_main["update_ms"] * 1e-3) / 1e9,
                                      "peak": peaks.get("hbm_gbs", 6650.0), "unit": "GB/s",
                                      "frac": kt_main["update_bytes"] / (kt_main["update_ms"] * 1e-3) / 1e9
                                              / peaks.get("hbm_gbs", 6650.0),
                                      "bytes_per_launch": kt_main["update_bytes"]}}
        # ---- the caller-side step in front of the kernel: fused 3-D RoPE + bf16 cast of Q, fp32 in
        aux = None
        try:
            import math
            from liteattention_b200.rope import rope_apply_bf16
            half = D // 2
            widths = [half - 2 * (half // 3), half // 3, half // 3]
            ang = torch.cat([torch.outer(torch.arange(1024, dtype=torch.float64),
                                         1.0 / torch.pow(10000.0, torch.arange(0, 2 * w_, 2, dtype=torch.float64) / (2 * w_)))
                             for w_ in widths], dim=1)
            freqs = torch.polar(torch.ones_like(ang), ang)
            gf = 21 if S == S_WAN else max(1, S // (45 * 80))
            grid_sizes = torch.tensor([[gf, 45, 80]] * B)
            xq = torch.randn(B, S, H, D, device=dev, dtype=torch.float32)
            for _ in range(3):
                rope_apply_bf16(xq, grid_sizes, freqs)
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0.record()
            for _ in range(10):
                rope_apply_bf16(xq, grid_sizes, freqs)
            r1.record()
            torch.cuda.synchronize()
            rms = r0.elapsed_time(r1) / 10
            rbytes = xq.numel() * 6.0
            aux = {"rope_cast_fp32_to_bf16": {"ms": rms, "bound": "hbm", "bytes_per_launch": rbytes,
                                              "achieved": rbytes / (rms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs", 6650.0),
                                              "unit": "GB/s", "frac": rbytes / (rms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0),
                                              "note": "one Q (or K) tensor of the workload; 4 B read + 2 B written per element"}}
            del xq
        except Exception as e:  # noqa: BLE001  (auxiliary line only; the headline numbers do not depend on it)
            aux = {"rope_cast_fp32_to_bf16": {"error": str(e)[:200]}}
        # the metric's other points (BASELINE.json: 0 / 42 / 77 %), kernel-level, same q/k/v -- and an UNBALANCED 42 % mask
        # (Bernoulli per tile: rows differ in length) next to the balanced one the headline uses
        sweep = []
        for sp in ((0.0, 0.21, 0.42, 0.57, 0.77) if args.sweep else (0.0, 0.77)):
            r = time_kernels(sp, max(5, args.steps // 2), 3)
            sweep.append({"sparsity": round(r["sparsity"], 4), "mask": "balanced (every row skips the same number of tiles)",
                          "fwd_ms": r["fwd_ms"], "update_ms": r["update_ms"],
                          "effective_tflops": dense_flops / (r["fwd_ms"] + r["update_ms"]) / 1e9,
                          "executed_tflops": r["exec_flops"] / r["fwd_ms"] / 1e9})
        rl_b, _ = synth.random_skip_list(B, H, qt, kt, args.sparsity, seed=99, device=dev)
        r = time_kernels(args.sparsity, max(5, args.steps // 2), 3, rl=rl_b)
        sweep.append({"sparsity": round(r["sparsity"], 4), "mask": "bernoulli per tile (rows of unequal length)",
                      "fwd_ms": r["fwd_ms"], "update_ms": r["update_ms"],
                      "effective_tflops": dense_flops / (r["fwd_ms"] + r["update_ms"]) / 1e9,
                      "executed_tflops": r["exec_flops"] / r["fwd_ms"] / 1e9})
        sweep.sort(key=lambda e: e["sparsity"])
        # update kernel with a must-do list (the README's text-token use case: the first 512 tokens are never skipped)
        md = LiteAttention._expand_must_do_list([511, 0], (B, H, qt, kt + 1), q, v)
        r = time_kernels(args.sparsity, max(5, args.steps // 2), 3, must_do=md)
        roofline["update_kernel"]["ms_with_must_do_list"] = r["update_ms"]
        del md
        # dense GPU comparators on the same box, same q/k/v (context, not the reference)
        comparators = None
        if world == 1 and not args.no_comparators:
            try:
                from baseline.comparators import time_dense_comparators
                comparators = time_dense_comparators(q, k, v, steps=max(3, min(args.steps, 5)), warmup=2)
                ok = {n: c for n, c in comparators.items() if "ms" in c}
                if ok:
                    best = min(ok, key=lambda n: ok[n]["ms"])
                    dense_ms = next(e["fwd_ms"] for e in sweep if e["sparsity"] == 0.0)
                    pts = sorted((e["sparsity"], e["fwd_ms"]) for e in sweep if e["mask"].startswith("balanced")) + \
                          [(kt_main["sparsity"], kt_main["fwd_ms"])]
                    pts.sort()
                    cross = None                      # sparsity at which la_fwd_kernel equals the best dense kernel (linear interp.)
                    for (s0, m0), (s1, m1) in zip(pts, pts[1:]):
                        if m0 >= ok[best]["ms"] >= m1 and m0 != m1:
                            cross = s0 + (s1 - s0) * (m0 - ok[best]["ms"]) / (m0 - m1)
                    comparators["summary"] = {"best_dense": best, "best_dense_ms": ok[best]["ms"], "la_fwd_dense_ms": dense_ms,
                                              "la_fwd_dense_over_best": dense_ms / ok[best]["ms"],
                                              "la_fwd_headline_ms": kt_main["fwd_ms"],
                                              "speedup_of_headline_over_best_dense": ok[best]["ms"] / kt_main["fwd_ms"],
                                              "crossover_sparsity": cross}
            except Exception as e:  # noqa: BLE001
                comparators = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
        cpu = None
        if not args.no_cpu and world == 1:
            c = cpu_sdpa_sample(S, 8, 1)
            cpu = {"value": c["tflops"], "unit": UNIT, "cores": c["cores"], "kind": "reference", "sample": c["sample"]}
        line = {
            "metric": METRIC, "value": dense_flops * world / (ms * 1e-3) / 1e12, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"Wan2.1-14B self-attention shape, B={B} per GPU, S={S}, H={H}, D={D}, bf16; "
                                   f"fixed random tile Skip-Mask at {step_sparsity:.1%} sparsity (128x176 tiles, runs of "
                                   "4), thr=-inf so the list is stationary; step = skip-list-gated forward + skip-list "
                                   "update" + ("" if world == 1 else "; batch-parallel, one prompt per GPU, O gathered to rank 0 ("
                                               + ("stored by the forward epilogue into rank 0's symmetric buffer over NVLink"
                                                  if bp.peer_store else f"NCCL gather pipelined in {n_groups} head groups") + ")"),
                       "batch_per_gpu": B, "seq_len": S, "heads": H, "head_dim": D, "sparsity": step_sparsity,
                       "parallelism": f"batch-parallel x{world}" + ("" if world == 1 else (", O stored by the forward epilogue into rank 0's symmetric buffer over NVLink (fused gather)" if bp.peer_store else ", NCCL gather" + ("" if args.gather == "nccl" else " (peer-store setup failed: " + str(bp.fallback_reason) + ")"))),
                       "l2": "q/k/v/o = 3.1 GB per step >> 126 MB L2, no flush needed"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "aux_kernels": aux,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line["sweep"] = sweep
        if comparators is not None:
            line["gpu_comparators"] = comparators
        if world > 1:
            line["gather_verified"] = gather_verified
            if seqpar is not None:
                line["seq_parallel"] = seqpar
            if e2e is not None:
                e2e["note"] = ("per rank: every rank copies ITS prompt host->device and ITS O device->host (what a DiT "
                               "pipeline does: each rank's O feeds its own next layer); not routed through the gather")
        line["derived_hopper_reference_ms"] = {"0%": 174, "42%": 104.5, "77%": 40.8,
                                               "note": "derived from README totals (BASELINE.md), H100-class, not measured"}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
