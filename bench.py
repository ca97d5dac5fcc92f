#!/usr/bin/env python
"""Headline benchmark: EVA attention forward, tokens/s at N=784 (28x28), C=192, h=3, d=64,
window 7, 49 landmarks (BASELINE.json config c3 -- one DeiT-tiny-p8 attention layer), fp16 I/O.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

One JSON line on stdout (rank 0).  `value` = attention-core tokens/s through the C ABI with q/k/v
resident in HBM; `e2e` = tokens/s of the drop-in module's forward(x) with x in pinned host memory
(H2D of x and D2H of y inside the timed region); `roofline` = algorithmic bytes of the core
(4*C*2 B/token) over the measured core time vs the measured HBM copy peak; `cpu_baseline` =
the CPU oracle port timed on this box's host cores on a bounded sample.

Multi-GPU (torchrun, one rank per GPU): the batch axis is sharded, no data-path collective
("weak" scaling); the timed region is bracketed by barriers and the max over ranks is reported.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'efficient-attention_b200'))
sys.path.insert(0, ROOT)

DIM, HEADS, GRID, WINDOW, LANDMARKS = 192, 3, 28, 7, 49
TOKENS = GRID * GRID
METRIC = 'EVA attn fwd tokens/sec at N=784,d=192'
WORKLOAD = 'c3: EVA layer fwd, N=784 (28x28), C=192, h=3, d=64, window 7, 49 landmarks, 2-D RPE, eval'


def ncu_traffic(batch):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this bench
    configuration (profiles/r01/ncu_traffic.json); None when the capture was taken at another batch size."""
    path = os.path.join(ROOT, 'profiles', 'r01', 'ncu_traffic.json')
    if not os.path.exists(path):
        return None
    d = json.load(open(path))
    return d['dram_bytes_read'] + d['dram_bytes_write'] if d.get('batch_per_gpu') == batch else None


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return float(json.load(open(path))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML every 2 ms WHILE the timed regions run (the recipe's clocks line).
    In-process NVML instead of an `nvidia-smi -lms` child: the child's start-up stalls kernel submission for tens of
    milliseconds, which is longer than a whole timed region."""

    REASONS = (('hw_slowdown', 0x8), ('hw_thermal_slowdown', 0x40), ('sw_thermal_slowdown', 0x20), ('sw_power_cap', 0x4))

    def __init__(self, index):
        self.index, self.rows, self.t, self.stop = index, [], None, threading.Event()
        self.max_mhz = None

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.t = None
        return self

    def _read(self):
        nv = self.nv
        while not self.stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), float(mhz), int(mask)))
            except Exception:
                pass
            self.stop.wait(0.002)

    def mark(self):
        """Samples taken before this call are warm-up and do not count."""
        self.t0 = time.perf_counter()

    def __exit__(self, *exc):
        self.stop.set()
        if self.t is not None:
            self.t.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows if r[0] >= getattr(self, 't0', 0.0)]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['unavailable']}
        reasons = [n for n, bit in self.REASONS if any(r[2] & bit for r in rows)]
        return {'sm_mhz': statistics.median(r[1] for r in rows), 'sm_max_mhz': self.max_mhz, 'reasons': reasons,
                'samples': len(rows)}


def build_layer(device, dtype):
    import warnings
    import efficient_attention as ea
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        m = ea.AttentionFactory.build_attention('eva', dict(
            dim=DIM, num_heads=HEADS, qkv_bias=True, attn_drop=0., proj_drop=0., fp32=False, use_rpe=True,
            window_size=WINDOW, attn_2d=True, overlap_window=False, adaptive_proj='default',
            num_landmarks=LANDMARKS, use_t5_rpe=False))
    # the reference init (std .02) makes every softmax uniform; use logits of order 1 instead
    g = torch.Generator().manual_seed(0)
    with torch.no_grad():
        for name, p in m.named_parameters():
            if p.dim() == 2 and 'bias_table' not in name:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / p.shape[1] ** 0.5))
            elif 'bias_table' in name:
                p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    return m.to(device=device, dtype=dtype).eval()


def cpu_baseline(seconds=12.0, batch=16):
    """The CPU oracle port (float32, all host threads) on a bounded sample of the same workload."""
    from oracle import eva_oracle as O
    m = build_layer('cpu', torch.float32)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = dict(num_heads=HEADS, window_size=WINDOW, attn_2d=True, overlap_window=False, adaptive_proj='default',
               num_landmarks=LANDMARKS, use_rpe=True, use_t5_rpe=False)
    torch.manual_seed(1)
    x = torch.randn(batch, GRID, GRID, DIM)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with torch.no_grad():
        for _ in range(2):
            O.eva_forward(sd, cfg, x)
        times = []
        t_end = time.perf_counter() + seconds
        while time.perf_counter() < t_end or len(times) < 3:
            t0 = time.perf_counter()
            O.eva_forward(sd, cfg, x)
            times.append(time.perf_counter() - t0)
    med = statistics.median(times)
    return {'value': batch * TOKENS / med, 'unit': 'tokens/s', 'cores': cores, 'kind': 'port',
            'sample': f'module forward(x), batch {batch} x {TOKENS} tokens, float32, {len(times)} passes, median'}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps, warm = max(args.steps, 1), args.warmup
    base = cpu_baseline(seconds=0.0, batch=16)  # warm-up + at least 3 passes
    from oracle import eva_oracle as O
    m = build_layer('cpu', torch.float32)
    sd = {k: v.detach() for k, v in m.state_dict().items()}
    cfg = dict(num_heads=HEADS, window_size=WINDOW, attn_2d=True, overlap_window=False, adaptive_proj='default',
               num_landmarks=LANDMARKS, use_rpe=True, use_t5_rpe=False)
    torch.manual_seed(1)
    batch = 16
    x = torch.randn(batch, GRID, GRID, DIM)
    with torch.no_grad():
        for _ in range(warm):
            O.eva_forward(sd, cfg, x)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.eva_forward(sd, cfg, x)
        dt = time.perf_counter() - t0
    val = batch * TOKENS * steps / dt
    base.update(value=val, sample=f'module forward(x), batch {batch} x {TOKENS} tokens per step, float32')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': 'tokens/s', 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'batch_per_step': batch,
                   'note': 'CPU oracle port of the reference PyTorch path (the Python reference cannot travel to the GPU box)'},
        'cpu_baseline': base,
        'e2e': {'value': val, 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))


def shard_range(total, world, rank):
    """[lo, hi) of the batch axis owned by `rank` (contiguous, covering, sizes differ by at most 1)."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_max_ms(ms, device, world):
    """A multi-GPU step is as slow as its slowest rank: MAX-reduce the device-timed milliseconds."""
    if world == 1:
        return float(ms)
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_tokens(per_rank_batch, world, uniform=True):
    """Whole-job tokens per step. Weak scaling: every rank processes `per_rank_batch` images."""
    if uniform or world == 1:
        return world * per_rank_batch * TOKENS
    import torch.distributed as dist
    t = torch.tensor([per_rank_batch * TOKENS], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def run_ours(args):
    import torch.distributed as dist
    from efficient_attention import _abi
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
        # several ranks stream from pinned host memory at once (e2e leg): keep this rank's threads -- and with them the pages
        # of its pinned buffers (first touch) -- on the NUMA node its GPU hangs off, instead of wherever torchrun started it
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = torch.cuda.get_device_properties(dev).uuid
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID('GPU-' + str(uuid))
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(uuid)).encode())
            pynvml.nvmlDeviceSetCpuAffinity(h)
        except Exception as e:                               # affinity is an optimisation, never a requirement
            if rank == 0:
                sys.stderr.write(f'bench: NUMA affinity not set ({type(e).__name__}: {e})\n')
    dtype = {'fp16': torch.float16, 'bf16': torch.bfloat16, 'fp32': torch.float32}[args.dtype]
    B, K, Wm = args.batch, args.steps, args.warmup
    layer = build_layer(dev, dtype)
    torch.manual_seed(1 + rank)
    x_host = torch.randn(B, GRID, GRID, DIM).to(dtype).pin_memory()
    x_dev = x_host.to(dev)
    elem = x_host.element_size()

    # ---- attention core through the C ABI, q/k/v resident in HBM ------------------------------
    with torch.no_grad():
        q, k, v, _ = layer._qkv_heads(x_dev.reshape(B, TOKENS, DIM))
        geom = _abi.eva_geometry(q, seq_shape=(GRID, GRID), window=WINDOW, ext=0, chunk=4, chunk_ext=0)
        ada, bias = layer._adaptive(), layer._local_bias().float().contiguous()

        def core():
            return _abi.eva_forward(q, k, v, geom, ada, bias=bias, return_path=True)

        clk = ClockSampler(local)
        clk.__enter__()                      # sampled over both timed regions (core and e2e)
        t_w, n_w = time.perf_counter(), 0
        while n_w < max(Wm, 3) or time.perf_counter() - t_w < 0.2:     # >= W launches and >= 0.2 s: clocks and caches settled
            _, path = core()
            n_w += 1
            if n_w % 16 == 0:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clk.mark()
        torch.cuda.synchronize()
        per_launch = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        ev0.record()
        for a_, b_ in per_launch:
            a_.record()
            core()
            b_.record()
        ev1.record()
        torch.cuda.synchronize()
        launch_ms = sum(a_.elapsed_time(b_) for a_, b_ in per_launch) / K     # device time of one eva_forward launch
        core_ms = reduce_max_ms(ev0.elapsed_time(ev1), dev, world)
        if world > 1:
            dist.barrier()

        # ---- end to end through the module's public forward(x), host buffers ----------------------
        from efficient_attention.streaming import HostPipeline
        # a serving loop with two host buffer pairs: batch i+1 is copied in while batch i is still being copied out (its
        # consumer reads y_hosts[i % 2] after pipe.done of that call); every step still moves its own x and y over PCIe
        x_hosts = [x_host, x_host.clone().pin_memory()]
        y_hosts = [torch.empty(B, GRID, GRID, DIM, dtype=dtype).pin_memory() for _ in range(2)]
        pipe = HostPipeline(layer, chunk=max(1, B // 8), defer_join=True)
        step_no = [0]

        def e2e_step():
            i = step_no[0] & 1
            step_no[0] += 1
            pipe(x_hosts[i], y_hosts[i])

        for _ in range(3):
            e2e_step()
        pipe.join()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        Ke = max(3, min(K, 10))
        e0.record()
        for _ in range(Ke):
            e2e_step()
        pipe.join()                              # the last batch's D2H copies are inside the timed region
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = reduce_max_ms(e0.elapsed_time(e1), dev, world)
        clk.__exit__(None, None, None)

    if rank == 0:
        tokens_per_step = aggregate_tokens(B, world)
        value = tokens_per_step * K / (core_ms * 1e-3)
        peak, peak_src = peaks()
        algo_bytes = 4 * DIM * elem * B * TOKENS           # read q,k,v once + write o once, per launch/rank
        achieved = algo_bytes / (launch_ms * 1e-3) / 1e9     # algorithmic bytes / average launch duration (CUDA events)
        out = {
            'metric': METRIC, 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': K, 'warmup': max(Wm, 3),
            'ms_per_step': core_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': args.dtype + ' I/O, f32 softmax/accumulate', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'batch_per_gpu': B, 'tokens_per_step': tokens_per_step,
                       'l2_policy': f'inputs larger than L2 (qkv {3 * DIM * elem * B * TOKENS / 2**20:.0f} MiB per GPU)',
                       'kernel_path': 'fused tcgen05/TMA' if path == 1 else 'generic two-stage CUDA-core',
                       'warmup_launches': n_w},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': ncu_traffic(B), 'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': algo_bytes, 'launch_ms': launch_ms,
                         'launches_timed': K,
                         'algorithmic_bytes_per_token': 4 * DIM * elem},
            'e2e': {'value': tokens_per_step * Ke / (e2e_ms * 1e-3), 'unit': 'tokens/s',
                    'h2d_bytes_per_step': x_host.numel() * elem, 'd2h_bytes_per_step': y_hosts[0].numel() * elem,
                    'steps': Ke, 'what': 'EVA.forward(x) via efficient_attention.streaming.HostPipeline: pinned-host x -> H2D, qkv Linear, '
                            'attention core, proj Linear, D2H -> pinned-host y, 8 chunks on 3 streams, two host buffer pairs (consecutive '
                            'batches overlap at their boundaries)'},
            'gpu_launches': K * 2,  # fused: weight-pack + fused kernel; generic: chunk_stats + window_attn
            'clocks': clk.summary(),
        }
        if world == 1 and not args.no_cpu:
            out['cpu_baseline'] = cpu_baseline()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=1024, help='images per GPU per step')
    ap.add_argument('--dtype', default='fp16', choices=['fp16', 'bf16', 'fp32'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
