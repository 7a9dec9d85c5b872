#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Metric: metric evals/sec, one eval = G^{-1}(z) + log det G(z) + grad_z log det G(z) at one latent
point, d=16, K=10,000 centroids (BASELINE.json configs[1]: N = 2^20 points per GPU).  One "step" is
one pass of the hot path over one batch of N synthetic points.  Under torchrun every rank
evaluates its own N points against replicated tables (weak scaling, no data-path collective,
SURVEY.md §8e); `value` is total evals over all ranks / max-over-ranks device time.

Extra keys: `roofline` (the fused metric + log-det tcgen05 kernel, tensor bound, timed alone against
MEASURED_PEAKS.json's bf16 burst rate; `roofline_gradient_kernel` is the same for the gradient kernel),
`cpu_baseline` (the CPU oracle port on a bounded sample, rank 0 / N=1 only), `e2e` (host-buffer
API: pinned H2D of z, evaluation, D2H of log det + grad, per step), `hmc` (config[2]: chain
leapfrog steps/s, 2^20 chains x 20 leapfrog), `flow` (config[3]: 65,536 sequences x 10 flow steps, metric
spectrum per step), `clocks`, `gpu_launches`.

`--impl reference` times the reference's own algorithm on the host cores (the CPU oracle port of
its eager PyTorch code -- the reference is Python and is not present on the GPU box).
"""
from __future__ import annotations

import argparse
import contextlib
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


def ncu_traffic(kernel_key):
    """dram bytes per launch from the committed `ncu --set full` capture (profiles/r1_traffic.json)."""
    p = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    try:
        return json.load(open(p)).get(kernel_key)
    except Exception:
        return None


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
def run_ours(args):
    import torch.distributed as dist
    from rlvae_b200 import MetricModel, MetricTensor, RiemannianHMCSampler, _capi
    from rlvae_b200.host_pipeline import HostEvaluator
    from rlvae_b200.synthetic import make_hmc_streams

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

    n = N_PER_GPU
    sm, z_cpu = tables_and_points(n, seed_offset=rank)
    mt = MetricTensor(D, device=dev)
    with contextlib.redirect_stdout(io.StringIO()):
        mt.load_pretrained(**sm.as_load_kwargs())
    if world > 1:   # tables are replicated: one broadcast at load
        for b in (mt.centroids, mt.metric_matrices):
            dist.broadcast(b, src=0)
    tab = mt._tables(dev)
    path_name = 'tensor' if (tab.tensor_capable and tab.tensor_auto) else 'direct'
    z = z_cpu.to(dev)
    out = {}
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)   # > 126 MB L2

    def step():
        nonlocal out
        out = mt.evaluate(z, want_ginv=True, want_g=False, want_logdet=True, want_grad=True, out=out)

    W, S = max(args.warmup, 3), max(args.steps, 1)
    for _ in range(W):
        step()
    barrier()
    _capi.lib().rlvae_launch_count(1)        # reset: count the kernels launched inside the timed region
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
    launches_timed = int(_capi.lib().rlvae_launch_count(0))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    dev_ms = max_over_ranks(dev_ms)
    ms_per_step = dev_ms / S
    value = world * n / (ms_per_step * 1e-3)

    # ---- the two tensor kernels of the step, each timed alone (CUDA events around the launch, L2 flushed)
    lib = _capi.lib()
    sym_tensor = path_name == 'tensor' and tab.symmetric
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
    # fused metric + log-det kernel: z -> packed G^-1, packed G, log det G (what rlvae_metric_eval runs first)
    def _ok(rc):
        assert rc == 0, lib.rlvae_last_error()

    fwd_ms = time_launch(lambda: _ok(lib.rlvae_metric_eval(tab.handle, _capi._ptr(z), n, None, None,
                                                           _capi._ptr(ld_buf), None, _capi._ptr(work),
                                                           _capi.PATH_AUTO, st)))
    # gradient kernel alone: contraction with the packed G the fused kernel left in the workspace
    g_src = out['g'] if 'g' in out and out['g'] is not None else None
    if g_src is None:
        g_src = mt.evaluate(z[: 1 << 16], want_ginv=False, want_g=True, want_logdet=False)['g'].repeat(n >> 16, 1, 1)
    grad_ms = time_launch(lambda: _ok(lib.rlvae_metric_grad(tab.handle, _capi._ptr(z), _capi._ptr(g_src), n,
                                                            -2.0 / float(tab.temperature) ** 2,
                                                            _capi._ptr(gr_buf), _capi.PATH_AUTO, st)))
    del g_src
    peaks, peak_src = measured_peaks()
    line = {}
    roof = roof_grad = None
    if rank == 0:
        tf32_rate = measure_tf32_peak(dev)
        bf16_rate = measure_bf16_rate(dev)
        peak = float(peaks['bf16_tflops'])          # burst figure: each kernel is timed alone
        if sym_tensor:
            # Tensor work in fp16-equivalent flops: a TF32 MAC costs two fp16 MACs of pipe time (the
            # measured TF32 GEMM rate is half the bf16 one).  ALGORITHMIC = symmetric tables, no padding:
            #   forward : 3 passes x 2K x (136 weighted-sum columns + 16 distance dims) [all fp16]
            #   gradient: 3 passes x 2K x (136 [fp16] + 16 [fp16] + 2 x 16 final contraction [TF32])
            f_fwd = n * 3 * 2 * K * (136 + 16)
            f_grad = n * 3 * 2 * K * (136 + 16 + 2 * 16)
            issued_fwd = n * 3 * 2 * tab.Kpad * (144 + 16)
            issued_grad = n * 3 * 2 * tab.Kpad * (144 + 16 + 2 * 16)
            note = ('fp16-equivalent tensor flops (TF32 MACs weighted x2: measured TF32 GEMM rate %.0f TF/s vs '
                    'bf16 %.0f TF/s in this run); symmetric tables -> 136 packed columns; dense-M definition '
                    'of SURVEY.md 8d would read %.0f TF/s fp32-equivalent' )
            ach = f_fwd / (fwd_ms * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'kernel': 'inverse_metric_h16_kernel (fused G^-1 + Cholesky log det)',
                    'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                    'traffic': ncu_traffic('inverse_metric_h16_kernel'), 'kernel_ms': fwd_ms,
                    'issued': issued_fwd / (fwd_ms * 1e-3) / 1e12,
                    'peak_source': f'MEASURED_PEAKS.json bf16_tflops burst ({peak_src})',
                    'same_run_cublas': {'bf16_tflops': bf16_rate, 'tf32_tflops': tf32_rate},
                    'note': note % (tf32_rate, bf16_rate, n * 2 * K * D * (D + 1) / (fwd_ms * 1e-3) / 1e12)}
            achg = f_grad / (grad_ms * 1e-3) / 1e12
            roof_grad = {'bound': 'tensor', 'kernel': 'metric_grad_h16_kernel', 'achieved': achg, 'peak': peak,
                         'unit': 'TFLOP/s', 'frac': achg / peak, 'traffic': ncu_traffic('metric_grad_h16_kernel'),
                         'kernel_ms': grad_ms, 'issued': issued_grad / (grad_ms * 1e-3) / 1e12}
        else:
            f_alg = n * 2 * K * D * (D + 1)
            ach = (3 if path_name == 'tensor' else 1) * f_alg / (fwd_ms * 1e-3) / 1e12
            roof = {'bound': 'tensor', 'kernel': 'inverse_metric_tc_kernel' if path_name == 'tensor'
                    else 'inverse_metric_direct_kernel', 'achieved': ach, 'peak': tf32_rate, 'unit': 'TFLOP/s',
                    'frac': ach / tf32_rate, 'traffic': None, 'kernel_ms': fwd_ms}

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
    del he

    # ---- HMC (BASELINE.json configs[2]): 2^20 chains x 20 leapfrog, one MCMC iteration
    hmc = None
    try:
        z0, gam, acc = make_hmc_streams(n, D, 1, seed=2 + rank)
        s = RiemannianHMCSampler(MetricModel(mt), mcmc_steps_nbr=1, n_lf=HMC_STEPS, eps_lf=0.03)
        z0, gam, acc = z0.to(dev), gam.to(dev), acc.to(dev)
        s.sample_with_streams(z0, gam, acc)                   # warm-up
        barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        s.sample_with_streams(z0, gam, acc)
        e1.record(); e1.synchronize()
        h_ms = max_over_ranks(e0.elapsed_time(e1))
        hmc = {'metric': 'hmc_chain_leapfrog_steps_per_sec', 'value': world * n * HMC_STEPS / (h_ms * 1e-3),
               'chains_per_gpu': n, 'n_lf': HMC_STEPS, 'mcmc_iterations': 1, 'ms': h_ms,
               'metric_evals_per_iteration': HMC_STEPS + 1}
    except Exception as e:   # never lose the headline because of the secondary measurement
        hmc = {'error': str(e)[:200]}

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
                                 'writes >2 GB of outputs'},
                'roofline': roof, 'roofline_gradient_kernel': roof_grad, 'cpu_baseline': cpu, 'e2e': e2e, 'hmc': hmc, 'flow': flow, 'clocks': clocks,
                'gpu_launches': launches_timed,
                'tflops_fp32_equiv': value * flops_per_eval(True) / 1e12}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
