#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric: metric evals/sec, one eval = G^{-1}(z) + log det G(z) + grad_z log det G(z) at one latent
point, d=16, K=10,000 centroids (BASELINE.json configs[1]: N = 2^20 points per GPU).  One "step" is
one pass of the hot path over one batch of N synthetic points.  Under torchrun every rank
evaluates its own N points against replicated tables (weak scaling, no data-path collective,
SURVEY.md §8e); `value` is total evals over all ranks / max-over-ranks device time.

Extra keys: `roofline` (the DOMINANT kernel of the step -- the gradient kernel -- timed with CUDA events on
the launching stream inside the timed steps, against MEASURED_PEAKS.json's sustained bf16 rate; three flop
conventions side by side; `roofline_forward_kernel` is the same for the fused metric + log-det kernel),
`cpu_baseline` (the CPU oracle port on a bounded sample, rank 0 / N=1 only), `e2e` (host-buffer API: pinned
H2D of z, evaluation, D2H of log det + grad, per step), `e2e_with_ginv` (G^-1 copied out too), `hmc`
(configs[2], weak: 2^20 chains per GPU x 20 leapfrog, one launch), `hmc_strong` (2^20 chains TOTAL over the
ranks + NCCL all_gather / all_reduce in the timed region), `small_T` (the step at T = 0.7), `d64` (configs[4]:
d = 64, K = 50k, 2^20 points per GPU), `flow` (configs[3]), `shard_check` (sharded == single-rank, bitwise),
`clocks`, `gpu_launches`.

`--impl reference` times the reference's own algorithm on the host cores (the CPU oracle port of
its eager PyTorch code -- the reference is Python and is not present on the GPU box).
"""
from __future__ import annotations

import argparse
import contextlib
import ctypes
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

D, K = 16, 10000
N_PER_GPU = 1 << 20
HMC_STEPS = 20
METRIC = 'metric_evals_per_sec'
UNIT = 'evals/s'


def flops_per_eval(with_grad=True):
    f = 2 * K * D * (D + 1)              # distance 2Kd + weighted sum 2Kd^2
    return 2 * f if with_grad else f


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.idx), '-lms', '50'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(',')]
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# --------------------------------------------------------------------------------------- helpers
def tables_and_points(n, seed_offset=0):
    from rlvae_b200.synthetic import make_points, make_synthetic_metric
    sm = make_synthetic_metric(K, D, seed=0)
    z = make_points(n, D, seed=1 + seed_offset)
    return sm, z


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        return json.load(open(p)), 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def measure_tf32_peak(dev):
    """cuBLAS TF32 dense GEMM, measured the way MEASURED_PEAKS.json measures bf16 (8192^3, best of 8)."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device=dev)
    b = torch.randn(8192, 8192, device=dev)
    best = 1e9
    for i in range(11):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); a @ b; e1.record(); e1.synchronize()
        if i >= 3:
            best = min(best, e0.elapsed_time(e1))
    torch.backends.cuda.matmul.allow_tf32 = old
    del a, b
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


def measure_bf16_rate(dev):
    """cuBLAS bf16 dense GEMM in THIS run (8192^3, best of 8): same-conditions cross-check of the
    MEASURED_PEAKS.json burst figure."""
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    best = 1e9
    for i in range(11):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); a @ b; e1.record(); e1.synchronize()
        if i >= 3:
            best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


def cpu_reference_rate(sm, budget_s=12.0, chunk=128, max_points=4096, threads=None):
    """The reference algorithm (CPU oracle port: [n,K,d,d] materialisation + LU inverse + slogdet +
    autograd) on the host cores: evals/s over a bounded sample of the same workload."""
    from oracle import metric_oracle as O
    from rlvae_b200.synthetic import make_points
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    t = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
    z = make_points(max_points, D, seed=1)
    O.eval_ginv_logdet_grad(z[:16], *t)                      # warm-up
    done, t0 = 0, time.perf_counter()
    while done < max_points:
        O.eval_ginv_logdet_grad(z[done:done + chunk], *t)
        done += chunk
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done, dt, threads


# --------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sm, _ = tables_and_points(16)
    W, S = max(args.warmup, 1), max(args.steps, 1)
    # each step = a bounded sample of the workload; keep the whole run within a few minutes
    per_step_budget = min(20.0, 150.0 / (W + S))
    rates, pts = [], 0
    for i in range(W + S):
        r, n, dt, th = cpu_reference_rate(sm, budget_s=per_step_budget, max_points=2048)
        if i >= W:
            rates.append((n, dt)); pts = n
    tot_n = sum(n for n, _ in rates); tot_t = sum(t for _, t in rates)
    value = tot_n / tot_t
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': S, 'warmup': W, 'ms_per_step': 1e3 * tot_t / S, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'G^-1 + log det G + grad_z log det G, d={D}, K={K}, N=2^20 per GPU '
                                   '(BASELINE.json configs[1])'},
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': th, 'kind': 'port',
                             'sample': f'{pts} points per step in 128-point chunks (the reference '
                                       'materialises [n,K,d,d]); torch CPU eager, all host threads'},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------- our arm
def committed_profile():
    """Numbers that can only come from an ncu capture (taken once per change, committed under
    profiles/): DRAM bytes per launch and tensor-pipe activity of the two kernels, for the SAME launches
    the timed step runs (profiles/README.md says which command produced them)."""
    for name in ('r2_traffic.json', 'r1_traffic.json'):
        p = os.path.join(ROOT, 'profiles', name)
        if os.path.isfile(p):
            try:
                d = json.load(open(p))
                d['_file'] = 'profiles/' + name
                return d
            except Exception:
                pass
    return {}


def run_ours(args):
    import torch.distributed as dist
    from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler, _capi
    from rlvae_b200.distributed import all_gather_rows, shard_bounds
    from rlvae_b200.host_pipeline import HostEvaluator
    from rlvae_b200.synthetic import make_hmc_streams, make_points

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def quiet_load(mt, **kw):
        with contextlib.redirect_stdout(io.StringIO()):
            mt.load_pretrained(**kw)

    n = N_PER_GPU
    sm, z_cpu = tables_and_points(n, seed_offset=rank)
    mt = MetricTensor(D, device=dev)
    quiet_load(mt, **sm.as_load_kwargs())
    if world > 1:   # tables are replicated: one broadcast at load
        for b in (mt.centroids, mt.metric_matrices):
            dist.broadcast(b, src=0)
    tab = mt._tables(dev)
    path_name = 'tensor' if (tab.tensor_capable and tab.tensor_auto) else 'direct'
    sym_tensor = path_name == 'tensor' and tab.symmetric
    z = z_cpu.to(dev)
    out = {}
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)   # > 126 MB L2
    lib = _capi.lib()

    def step():
        nonlocal out
        out = mt.evaluate(z, want_ginv=True, want_g=False, want_logdet=True, want_grad=True, out=out)

    W, S = max(args.warmup, 3), max(args.steps, 1)
    for _ in range(W):
        step()
    barrier()
    lib.rlvae_launch_count(1)        # reset: count the kernels launched inside the timed region
    lib.rlvae_profile_begin(S)       # CUDA events around each kernel of every timed step, on the launching stream
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    t_wall0 = time.time()
    barrier()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(S)]
    for i in range(S):
        flush.zero_()                      # L2 flush between timed iterations (outside the event pair)
        ev[i][0].record()
        step()
        ev[i][1].record()
    barrier()
    launches_timed = int(lib.rlvae_launch_count(0))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / S
    value = world * n / (ms_per_step * 1e-3)
    # the kernels of the timed steps themselves (what the roofline block reports)
    in_step = None
    if int(lib.rlvae_profile_count()) == S:
        acc3 = [0.0, 0.0, 0.0]
        buf = (ctypes.c_float * 3)()
        for i in range(S):
            assert lib.rlvae_profile_read(i, buf) == 0, lib.rlvae_last_error()
            for j in range(3):
                acc3[j] += float(buf[j])
        in_step = {'forward_ms': acc3[0] / S, 'fallback_pass_ms': acc3[1] / S, 'gradient_ms': acc3[2] / S}
    lib.rlvae_profile_end()

    # ---- each tensor kernel also timed ALONE (L2 flushed, idle GPU before it: burst clocks)
    work = torch.empty(int(lib.rlvae_metric_eval_workspace(n, D)) // 4, device=dev, dtype=torch.float32)
    ld_buf = torch.empty(n, device=dev)
    gr_buf = torch.empty(n, D, device=dev)

    def time_launch(fn, reps=5):
        ts = []
        for i in range(3 + reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        return sum(ts) / len(ts)

    st = _capi._stream(z)

    def _ok(rc):
        assert rc == 0, lib.rlvae_last_error()

    # the SAME forward launch the step runs: expanded G^-1 + log det + packed G (for the gradient kernel);
    # measured as (forward + gradient call) - (gradient alone) would mix kernels, so it is taken from the
    # profile events of a call that also runs the gradient
    fwd_alone_ms = grad_alone_ms = None
    if sym_tensor:
        lib.rlvae_profile_begin(8)
        buf = (ctypes.c_float * 3)()
        fa, ga = [], []
        for i in range(8):
            flush.zero_()
            torch.cuda.synchronize(dev)
            _ok(lib.rlvae_metric_eval(tab.handle, _capi._ptr(z), n, _capi._ptr(out['ginv']), None, _capi._ptr(ld_buf),
                                      _capi._ptr(gr_buf), _capi._ptr(work), _capi.PATH_AUTO, st))
            _ok(lib.rlvae_profile_read(i, buf))
            if i >= 3:
                fa.append(float(buf[0])); ga.append(float(buf[2]))
        lib.rlvae_profile_end()
        fwd_alone_ms, grad_alone_ms = sum(fa) / len(fa), sum(ga) / len(ga)
    else:
        fwd_alone_ms = time_launch(lambda: _ok(lib.rlvae_metric_eval(tab.handle, _capi._ptr(z), n, _capi._ptr(out['ginv']),
                                                                      None, _capi._ptr(ld_buf), None, _capi._ptr(work),
                                                                      _capi.PATH_AUTO, st)))
    peaks, peak_src = measured_peaks()
    line = {}
    roof = roof_fwd = None
    if rank == 0:
        tf32_rate = measure_tf32_peak(dev)
        bf16_rate = measure_bf16_rate(dev)
        prof = committed_profile()
        burst, sustained = float(peaks['bf16_tflops']), float(peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']))
        if sym_tensor and in_step is not None:
            # Tensor work in fp16-equivalent flops: a TF32 MAC costs two fp16 MACs of pipe time (the measured
            # TF32 GEMM rate is half the bf16 one).  Three conventions, printed side by side:
            #   packed   ALGORITHMIC work of this design: symmetric tables -> 136 packed columns (+16 distance
            #            dims), x3 split passes, no padding          
            #   dense_8d SURVEY.md 8d's dense-M count 2Kd(d+1) per eval and kernel, x3 split passes
            #            (exceeds the packed count by 272/152: symmetry is not credited)
            #   fp32_eq  the same dense count with NO pass multiplier ("fp32-equivalent" flops)
            dense = n * 2.0 * K * D * (D + 1)

            def block(kernel, ms_step, ms_alone, packed_flops, issued_flops, key):
                ach = packed_flops / (ms_step * 1e-3) / 1e12
                pk = prof.get(key) if isinstance(prof.get(key), dict) else {}
                return {'bound': 'tensor', 'kernel': kernel, 'achieved': ach, 'peak': burst, 'unit': 'TFLOP/s',
                        'frac': ach / burst,
                        'frac_of_sustained_peak': ach / sustained,
                        'frac_dense_8d': 3 * dense / (ms_step * 1e-3) / 1e12 / burst,
                        'frac_fp32_equiv': dense / (ms_step * 1e-3) / 1e12 / burst,
                        'issued_tflops': issued_flops / (ms_step * 1e-3) / 1e12,
                        'kernel_ms': ms_step, 'kernel_ms_after_idle': ms_alone,
                        'tensor_pipe_active': pk.get('tensor_pipe_active_pct'),
                        'traffic': pk.get('dram_bytes'), 'traffic_algorithmic': pk.get('algorithmic_bytes'),
                        'ncu_source': prof.get('_file'),
                        'timing': 'kernel_ms = CUDA events on the launching stream around this kernel inside the '
                                  'timed steps (average over the steps).  peak = the BURST bf16 figure (the '
                                  'conservative denominator: the steps are 16 ms long and run against the 1 kW power '
                                  'cap, frac_of_sustained_peak uses the 4-second back-to-back figure); '
                                  'kernel_ms_after_idle = the same launch right after a device synchronisation '
                                  '(clocks still ramping: slower, reported for completeness)',
                        'peak_source': f'MEASURED_PEAKS.json bf16_tflops (burst) / bf16_tflops_sustained ({peak_src})'}

            f_fwd = n * 3.0 * 2 * K * (136 + 16)
            f_grad = n * 3.0 * 2 * K * (136 + 16 + 2 * 16)
            issued_fwd = n * 3.0 * 2 * tab.Kpad * (144 + 16)
            issued_grad = n * 3.0 * 2 * tab.Kpad * (144 + 16 + 2 * 16)
            roof_g = block('metric_grad_h16_kernel', in_step['gradient_ms'], grad_alone_ms, f_grad, issued_grad,
                           'metric_grad_h16_kernel')
            roof_fwd = block('inverse_metric_h16_kernel (fused G^-1 + Cholesky log det + packed G)',
                             in_step['forward_ms'], fwd_alone_ms, f_fwd, issued_fwd, 'inverse_metric_h16_kernel')
            roof = roof_g if in_step['gradient_ms'] >= in_step['forward_ms'] else roof_fwd     # the DOMINANT kernel
            roof = dict(roof)
            roof['dominant'] = roof['kernel']
            roof['step_share'] = {'forward_ms': in_step['forward_ms'], 'fallback_pass_ms': in_step['fallback_pass_ms'],
                                  'gradient_ms': in_step['gradient_ms'], 'step_ms': ms_per_step}
            roof['same_run_cublas'] = {'bf16_tflops': bf16_rate, 'tf32_tflops': tf32_rate}
            roof['step_frac'] = (f_fwd + f_grad) / (ms_per_step * 1e-3) / 1e12 / burst
        else:
            f_alg = n * 2.0 * K * D * (D + 1)
            ach = (3 if path_name == 'tensor' else 1) * f_alg / (fwd_alone_ms * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'kernel': 'inverse_metric_tc_kernel' if path_name == 'tensor'
                    else 'inverse_metric_direct_kernel', 'achieved': ach, 'peak': tf32_rate, 'unit': 'TFLOP/s',
                    'frac': ach / tf32_rate, 'traffic': None, 'kernel_ms': fwd_alone_ms}

    # ---- shard check: the gathered sharded evaluation equals the single-rank evaluation (2^16 points)
    shard_check = None
    try:
        nz = 1 << 16
        zg = make_points(nz, D, seed=77).to(dev)
        lo, hi = shard_bounds(nz, world, rank)
        mine = mt.evaluate(zg[lo:hi].contiguous(), want_ginv=True, want_logdet=True, want_grad=True)
        gath = {k: all_gather_rows(mine[k], nz) for k in ('ginv', 'logdet_g', 'grad_logdet_g')}
        full = mt.evaluate(zg, want_ginv=True, want_logdet=True, want_grad=True)
        diffs = {k: float((gath[k] - full[k]).abs().max().item()) for k in gath}
        ok = all(torch.equal(gath[k], full[k]) for k in gath)
        okt = torch.tensor([1.0 if ok else 0.0], device=dev)
        if world > 1:
            dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        shard_check = {'points': nz, 'world': world, 'bitwise_equal': bool(okt.item() == 1.0), 'max_abs_diff': diffs}
        assert shard_check['bitwise_equal'], f'sharded evaluation differs from the single-rank one: {diffs}'
        del zg, mine, gath, full
    except AssertionError:
        raise
    except Exception as e:
        shard_check = {'error': str(e)[:200]}

    # ---- end to end through the host-buffer API (pinned H2D of z, D2H of log det + grad)
    he = HostEvaluator(mt, chunk=1 << 17, want_grad=True)
    z_pin = z_cpu.pin_memory()
    ld_pin = torch.empty(n).pin_memory()
    gr_pin = torch.empty(n, D).pin_memory()
    for _ in range(2):
        io_bytes = he(z_pin, ld_pin, gr_pin)
    barrier()
    e2e_steps = max(2, min(S, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        io_bytes = he(z_pin, ld_pin, gr_pin)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / e2e_steps
    e2e = {'value': world * n / e2e_s, 'unit': UNIT, 'h2d_bytes_per_step': io_bytes['h2d_bytes'],
           'd2h_bytes_per_step': io_bytes['d2h_bytes'], 'ms_per_step': 1e3 * e2e_s,
           'api': 'rlvae_b200.host_pipeline.HostEvaluator (pinned host z in, log det + grad out; '
                  'G^-1 stays on device)'}
    # the same API used as a stream: one batch in flight, batch i is read back while batch i + 1 runs (every step still
    # copies its inputs H2D and its results D2H inside the timed region; `e2e` above stays the synchronous per-step number)
    e2e_pipe = None
    try:
        ld2, gr2 = torch.empty(n).pin_memory(), torch.empty(n, D).pin_memory()
        bufs = ((ld_pin, gr_pin), (ld2, gr2))
        barrier()
        t0 = time.perf_counter()
        pending = None
        for i in range(e2e_steps):
            ev, _ = he.submit(z_pin, bufs[i & 1][0], bufs[i & 1][1])
            if pending is not None:
                he.wait(pending)
            pending = ev
        he.wait(pending)
        barrier()
        sp = max_over_ranks(time.perf_counter() - t0) / e2e_steps
        e2e_pipe = {'value': world * n / sp, 'unit': UNIT, 'ms_per_step': 1e3 * sp,
                    'h2d_bytes_per_step': io_bytes['h2d_bytes'], 'd2h_bytes_per_step': io_bytes['d2h_bytes'],
                    'same_results_as_e2e': bool(torch.equal(ld_pin, ld2) and torch.equal(gr_pin, gr2)),
                    'api': 'HostEvaluator.submit / wait: one batch in flight, results of batch i read while batch i+1 runs'}
        del ld2, gr2
    except Exception as e:
        e2e_pipe = {'error': str(e)[:200]}
    # the same with G^-1 copied out as well (1 KB per point: PCIe bound)
    e2e_ginv = None
    try:
        gi_pin = torch.empty(n, D, D).pin_memory()
        io2 = he(z_pin, ld_pin, gr_pin, ginv_host=gi_pin)
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            io2 = he(z_pin, ld_pin, gr_pin, ginv_host=gi_pin)
        barrier()
        s2 = max_over_ranks(time.perf_counter() - t0) / 2
        e2e_ginv = {'value': world * n / s2, 'unit': UNIT, 'h2d_bytes_per_step': io2['h2d_bytes'],
                    'd2h_bytes_per_step': io2['d2h_bytes'], 'ms_per_step': 1e3 * s2,
                    'd2h_gbs_per_gpu': io2['d2h_bytes'] / s2 / 1e9,
                    'api': 'HostEvaluator(..., ginv_host=pinned [N,16,16]): every output of the metric copied to the host'}
        del gi_pin
    except Exception as e:
        e2e_ginv = {'error': str(e)[:200]}
    del he

    # ---- HMC (BASELINE.json configs[2]): 20 leapfrog steps, one MCMC iteration = ONE kernel launch
    def time_hmc(nch, seed, gather):
        z0, gam, acc = make_hmc_streams(nch, D, 1, seed=seed)
        if gather:      # strong scaling: rows of the GLOBAL draws that belong to this rank
            lo, hi = shard_bounds(nch, world, rank)
            z0, gam, acc = z0[lo:hi].contiguous(), gam[:, lo:hi].contiguous(), acc[:, lo:hi].contiguous()
        s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=1, n_lf=HMC_STEPS, eps_lf=0.03)
        z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
        best = None
        for rep in range(3):
            barrier()
            lib.rlvae_launch_count(1)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            rec = {}
            zf = s.sample_with_streams(z0, gam, acc, record=rec if gather else None)
            if gather:   # what north_star names: gather the samples, reduce the scalar (accept count)
                zall = all_gather_rows(zf, nch)
                cnt = rec['moves'][0].sum()
                if world > 1:
                    dist.all_reduce(cnt)
            e1.record(); e1.synchronize()
            ms = max_over_ranks(e0.elapsed_time(e1))
            if rep and (best is None or ms < best[0]):
                best = (ms, int(lib.rlvae_launch_count(0)), float(cnt.item()) / nch if gather else None,
                        tuple(zall.shape) if gather else None)
        return best

    hmc = hmc_strong = None
    try:
        h_ms, h_launch, _, _ = time_hmc(n, 2 + rank, False)
        hmc = {'metric': 'hmc_chain_leapfrog_steps_per_sec', 'value': world * n * HMC_STEPS / (h_ms * 1e-3),
               'chains_per_gpu': n, 'n_lf': HMC_STEPS, 'mcmc_iterations': 1, 'ms': h_ms,
               'metric_evals_per_iteration': HMC_STEPS + 1, 'kernel_launches_per_iteration': h_launch,
               'fused_trajectory_kernel': bool(_capi.hmc_fused_available(tab)), 'scaling': 'weak'}
    except Exception as e:   # never lose the headline because of the secondary measurement
        hmc = {'error': str(e)[:200]}
    try:
        s_ms, s_launch, acc_rate, gshape = time_hmc(n, 2, True)
        hmc_strong = {'metric': 'hmc_chain_leapfrog_steps_per_sec', 'value': n * HMC_STEPS / (s_ms * 1e-3),
                      'chains_total': n, 'chains_per_gpu': n // world, 'n_lf': HMC_STEPS, 'ms': s_ms,
                      'scaling': 'strong', 'accept_rate': acc_rate, 'gathered_shape': list(gshape),
                      'collectives_in_timed_region': 'all_gather(z) + all_reduce(accept count) over NCCL'
                      if world > 1 else 'none (1 rank)'}
    except Exception as e:
        hmc_strong = {'error': str(e)[:200]}

    # ---- small temperature (the reference's T = 0.7, conf/model/hybrid_rlvae.yaml:41): the same step
    small_t = None
    try:
        mt7 = MetricTensor(D, device=dev)
        kw = sm.as_load_kwargs(); kw['temperature'] = 0.7
        quiet_load(mt7, **kw)
        # half of the points next to centroids (at T = 0.7 a random point sees no centroid at all)
        z7 = torch.cat([z[: n // 2], mt7.centroids[torch.arange(n // 2, device=dev) % K] + 0.05 * z[n // 2:]]).contiguous()
        o7 = {}
        for _ in range(2):
            o7 = mt7.evaluate(z7, want_ginv=True, want_logdet=True, want_grad=True, out=o7)
        barrier()
        ts = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            o7 = mt7.evaluate(z7, want_ginv=True, want_logdet=True, want_grad=True, out=o7)
            e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        t7 = max_over_ranks(sum(ts) / len(ts))
        small_t = {'temperature': 0.7, 'ms_per_step': t7, 'value': world * n / (t7 * 1e-3), 'unit': UNIT,
                   'implementation': mt7.kernel_info().get('implementation'),
                   'finite': bool(torch.isfinite(o7['logdet_g']).all().item())}
        del mt7, z7, o7
    except Exception as e:
        small_t = {'error': str(e)[:200]}

    # ---- FlowManager temporal flow (BASELINE.json configs[3]): B = 65,536 sequences x 10 timesteps,
    # flows in stock torch, then ONE fused metric evaluation (log det G + spectrum) over [B*T, d]
    flow = None
    try:
        from rlvae_b200 import FlowManager
        torch.manual_seed(7)
        fm = FlowManager(latent_dim=D, n_flows=8, device=dev).to(dev).eval()
        zf0 = torch.randn(65536, D, device=dev)
        with torch.no_grad():
            fm.metric_along_flow(mt, zf0, n_obs=10, want_spectrum=True)      # warm-up
            barrier()
            e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            z_seq, _ = fm.apply_flows([zf0], n_obs=10)
            e1.record()
            zz = torch.stack(z_seq, dim=1).reshape(-1, D).contiguous()
            sp = mt.compute_metric_spectrum(zz)
            e2.record(); e2.synchronize()
        f_ms, m_ms = max_over_ranks(e0.elapsed_time(e1)), max_over_ranks(e1.elapsed_time(e2))
        flow = {'sequences_per_gpu': 65536, 'timesteps': 10, 'apply_flows_ms': f_ms,
                'metric_spectrum_ms': m_ms, 'metric_evals_per_sec': world * 655360 / (m_ms * 1e-3),
                'finite': bool(torch.isfinite(sp['logdet_G']).all().item())}
        del fm, zf0, z_seq, zz, sp
    except Exception as e:
        flow = {'error': str(e)[:200]}

    # ---- pythae variant: log_pi + grad log_pi of RHVAESampler at one position per point
    # (rlvae_pythae_eval: forward kernel + unit-weight gradient kernel + finish), and the reference-sized use:
    # OfficialRHVAESampler.sample_prior(32) = 100 MCMC x 15 leapfrog steps at the hard-coded T = 0.1
    pythae = None
    try:
        from rlvae_b200 import MetricModel, OfficialRHVAESampler, _capi
        tab = mt._tables(dev)
        with torch.no_grad():
            _capi.pythae_eval(tab, z)
            barrier()
            ts = []
            for _ in range(3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                pg, pl, ps = _capi.pythae_eval(tab, z)
                e1.record(); e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            tp = max_over_ranks(sum(ts) / len(ts))
            osamp = OfficialRHVAESampler(MetricModel(mt))
            osamp.sample_prior(32); torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            zo = osamp.sample_prior(32); torch.cuda.synchronize(dev)
            t_off = max_over_ranks((time.perf_counter() - t0) * 1e3)
            # 1024 prior samples = 32 of the reference's 32-chain batches, drawn in its order, one library call
            t0 = time.perf_counter()
            zo2 = osamp.sample_prior(1024); torch.cuda.synchronize(dev)
            t_off2 = max_over_ranks((time.perf_counter() - t0) * 1e3)
        pythae = {'what': 'log|det G^-1| + (1/T^2) G^T sum_k w_k M_k^T (c_k - z), 2^20 points per GPU', 'ms': tp,
                  'value': world * n / (tp * 1e-3), 'unit': UNIT,
                  'launches': 'forward kernel (+ its normally-empty Cholesky fallback pass), unit-weight pass of the '
                              'gradient kernel, finish kernel with the per-row error bound, per-centroid kernel over '
                              'the flagged rows (none at this configuration)',
                  'finite': bool(torch.isfinite(pg).all().item() and torch.isfinite(pl).all().item()),
                  'official_sample_prior_32': {'mcmc_steps': 100, 'n_lf': 15, 'temperature': 0.1, 'wall_ms': t_off,
                                               'finite': bool(torch.isfinite(zo).all().item())},
                  'official_sample_prior_1024': {'batches_of_32_merged': 32, 'wall_ms': t_off2,
                                                 'finite': bool(torch.isfinite(zo2).all().item())}}
        del pg, pl, ps, zo, zo2, osamp
    except Exception as e:
        pythae = {'error': str(e)[:200]}

    # ---- large-metric stress (BASELINE.json configs[4]): d = 64, K = 50,000, N = 2^20 points per GPU
    # (8M over 8 GPUs), tables generated on the device from a fixed seed, evaluated in 2^17-point chunks
    d64 = None
    del out, work, ld_buf, gr_buf, z
    torch.cuda.empty_cache()
    try:
        d64 = bench_d64(dev, world, rank, barrier, max_over_ranks)
    except Exception as e:
        d64 = {'error': str(e)[:300]}

    if rank == 0:
        cpu = None
        if world == 1:
            r, npts, dt, th = cpu_reference_rate(sm, budget_s=12.0)
            cpu = {'value': r, 'unit': UNIT, 'cores': th, 'kind': 'port',
                   'sample': f'{npts} points of the same workload in 128-point chunks, {dt:.1f} s; CPU oracle '
                             'port of the reference eager PyTorch path (G^-1, log det via inv+slogdet, '
                             'grad by autograd)'}
            # the same for the sampler: RiemannianHMCSampler.sample on the host cores
            try:
                from oracle import metric_oracle as O
                nch = 16
                hz0, hgam, hacc = make_hmc_streams(nch, D, 1, seed=2)
                tt = (sm.centroids, sm.metric_matrices, sm.temperature, sm.regularization)
                t0 = time.perf_counter()
                O.hmc_sample(tt, hz0, hgam, hacc, HMC_STEPS, 0.03)
                hdt = time.perf_counter() - t0
                if isinstance(hmc, dict):
                    hmc['cpu_baseline'] = {'value': nch * HMC_STEPS / hdt, 'unit': 'chain-leapfrog-steps/s',
                                           'cores': th, 'kind': 'port',
                                           'sample': f'{nch} chains x {HMC_STEPS} leapfrog x 1 MCMC iteration, K={K}, '
                                                     f'{hdt:.1f} s (the reference evaluates the metric 2*n_lf+2 times)'}
            except Exception as e:
                if isinstance(hmc, dict):
                    hmc['cpu_baseline'] = {'error': str(e)[:200]}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': S, 'warmup': W,
                'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32 (split-fp16 / 3xTF32 tensor products with fp32 accumulate: fp32-level accuracy)' if path_name == 'tensor' else 'f32',
                'data': 'synthetic',
                'config': {'workload': f'G^-1 + log det G + grad_z log det G, d={D}, K={K}, N=2^20 per GPU '
                                       '(BASELINE.json configs[1])', 'points_per_gpu': n, 'path': path_name, 'implementation': mt.kernel_info().get('implementation'),
                           'parallelism': f'points sharded over {world} GPU(s), tables replicated',
                           'l2': 'L2 flushed (256 MB write) before every timed step; each step also '
                                 'writes >1.6 GB of outputs'},
                'roofline': roof, 'roofline_forward_kernel': roof_fwd, 'cpu_baseline': cpu, 'e2e': e2e,
                'e2e_pipelined': e2e_pipe, 'e2e_with_ginv': e2e_ginv, 'hmc': hmc, 'hmc_strong': hmc_strong, 'small_T': small_t, 'd64': d64,
                'flow': flow, 'pythae': pythae, 'shard_check': shard_check, 'clocks': clocks,
                'gpu_launches': launches_timed,
                'tflops_fp32_equiv': value * flops_per_eval(True) / 1e12}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def bench_d64(dev, world, rank, barrier, max_over_ranks, n=1 << 20, K64=50000, chunk=1 << 17):
    """BASELINE.json configs[4]: d = 64, K = 50,000 centroids, 2^20 points per GPU (8M over 8 GPUs).  Tables
    and points from a fixed device seed (the recipe of scripts/stress_d64.py: M = L L^T, rescaled so that
    G^-1 is O(1)); one step = G^-1 + log det G (+ grad_z log det G) for all points, in chunks."""
    from rlvae_b200 import MetricTensor
    d = 64
    g = torch.Generator(device=dev).manual_seed(1234)
    c = torch.randn(K64, d, device=dev, generator=g)
    L = torch.tril(torch.randn(K64, d, d, device=dev, generator=g)) * d ** -0.5
    M = L @ L.transpose(1, 2)
    del L
    M = 0.5 * (M + M.transpose(1, 2))
    T, lam = 0.75 * d ** 0.5, 0.01
    gz = torch.Generator(device=dev).manual_seed(4321 + rank)
    z = torch.randn(n, d, device=dev, generator=gz)
    w = torch.exp(-torch.cdist(z[:64].double(), c.double()) ** 2 / T ** 2).sum(1).mean().item()
    M = (M / w).contiguous()
    mt = MetricTensor(d, device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(c, M, temperature=T, regularization=lam)
    del M
    info = mt.kernel_info()
    out = {}

    def run(want_grad, npts):
        nonlocal out
        for lo in range(0, npts, chunk):
            zz = z[lo:min(lo + chunk, npts)]
            if out and out['ginv'].shape[0] != zz.shape[0]:
                out = {}
            out = mt.evaluate(zz, want_ginv=True, want_logdet=True, want_grad=want_grad, out=out)

    res = {'latent_dim': d, 'n_centroids': K64, 'points_per_gpu': n, 'chunk_points': chunk,
           'implementation': info.get('implementation'), 'tables_bytes_fp32': K64 * (d * d + d) * 4}
    run(False, chunk)                       # warm-up (one chunk)
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); run(False, n); e1.record(); e1.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1))
    res['forward'] = {'what': 'G^-1 + log det G', 'ms': ms, 'value': world * n / (ms * 1e-3), 'unit': UNIT,
                      'tflops_fp32_equiv_dense': world * n * 2.0 * K64 * d * (d + 1) / (ms * 1e-3) / 1e12}
    res['finite'] = bool(torch.isfinite(out['logdet_g']).all().item())
    # with the gradient: full size when the gradient runs on the tensor cores, a bounded sample otherwise
    tensor_grad = 'gradient: partial tiles' in str(info.get('implementation'))
    ng = n if tensor_grad else 4096
    run(True, min(ng, chunk))
    barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); run(True, ng); e1.record(); e1.synchronize()
    msg = max_over_ranks(e0.elapsed_time(e1))
    res['with_grad'] = {'what': 'G^-1 + log det G + grad_z log det G', 'points_per_gpu': ng, 'ms': msg,
                        'value': world * ng / (msg * 1e-3), 'unit': UNIT, 'gradient_on_tensor_cores': tensor_grad}
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
