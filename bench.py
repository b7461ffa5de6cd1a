#!/usr/bin/env python
"""bench.py -- signal samples/sec basecalled (sup@v3.3 XNA) on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp16|bf16]

A "step" is one pass of the hot path over one batch: BASELINE.json configs[1] = sup@v3.3 architecture,
UB "X" alphabet (n_base 5), batch 512 x 4000-sample chunks: conv stem -> 5 LSTMs -> CRF head -> CRF
posteriors -> max-marginal Viterbi -> left-packed base strings.  Metric (bonito/cli/basecaller.py:153-161):
input signal samples consumed per second.

  value   whole-job samples/s with the batch already resident in HBM (device-timed, max over ranks)
  e2e     same metric through the host-buffer C-ABI calls xb_compute_scores_submit / _wait (the pipelined form of
          xb_compute_scores_host that crf.basecall() uses): pinned host signal in, H2D + encoder + decode + D2H of packed
          strings of every step inside the timed region, batch i+1 submitted before batch i is collected
  roofline  dominant kernel (LSTM recurrence, tensor-core bound): algorithmic FLOPs / CUDA-event time
  cpu_baseline  the oracle's restatement of the reference CPU path, timed on this box's host cores on a
          bounded sample (rank 0, N=1 only)

--impl reference times the reference's own CPU implementation of the path (its torch modules as restated
in oracle/bonito_oracle.py + the C CRF decode) on the host cores, same metric.
Multi-GPU (torchrun): each rank runs the same batch shape on its own GPU (reads/chunks shard with no
data-path collective -> weak scaling); NCCL only gathers the timing / counters.
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

ALPHABET = ['N', 'A', 'C', 'G', 'T', 'X']
N_BASE, STATE_LEN, FEATURES = 5, 3, 768
CHUNK, BATCH = 4000, 512
T_STEPS = CHUNK // 5
FLOP_PER_CHUNK = {   # SURVEY.md 8(d), 2*MAC
    'conv': 2 * (4 * 1 * 5 * 4000 + 16 * 4 * 5 * 4000 + 768 * 16 * 19 * 800),
    'lstm': 5 * 800 * 2 * (2 * 3072 * 768),
    'head': 2 * 768 * 625 * 800,
}


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tf_burst': p['bf16_tflops'], 'tf_sustained': p['bf16_tflops_sustained'],
                'source': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        try:                                            # NVML: a sample every 20 ms (nvidia-smi needs ~50 ms per call)
            import pynvml
            pynvml.nvmlInit()
            dev = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(dev, pynvml.NVML_CLOCK_SM)
            bits = {'hw_slowdown': pynvml.nvmlClocksThrottleReasonHwSlowdown,
                    'hw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    'sw_thermal_slowdown': pynvml.nvmlClocksThrottleReasonSwThermalSlowdown,
                    'sw_power_cap': pynvml.nvmlClocksThrottleReasonSwPowerCap}
            names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
            while not self.stop_flag.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(dev, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(dev)
                watts = pynvml.nvmlDeviceGetPowerUsage(dev) / 1000.0
                self.rows.append([str(sm), str(mx), str(watts)] + ['Active' if reasons & bits[n] else 'Not Active' for n in names])
                self.stop_flag.wait(0.02)
            return
        except Exception:
            pass
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx or None, 'reasons': sorted(reasons),
                'samples': len(sm)}


def workload_config(N, L):
    return {'workload': 'configs[1]: sup@v3.3 UB X (n_base 5), batch %d x %d-sample chunks per GPU, conv stem + '
                        '5 LSTM + CRF head + posteriors + Viterbi + left-pack; random-init weights; '
                        'per-step working set (activations 0.6 GB, gates 2.5 GB, scores 1.2 GB) exceeds the 126 MB L2, '
                        'no explicit flush' % (N, L),
            'batch_per_gpu': N, 'chunk': L, 'n_base': N_BASE, 'sharding': 'chunks across GPUs, no collective'}


def cpu_reference_pass(n_chunks, threads, sd=None, repeats=1):
    """One pass of the reference CPU path over n_chunks chunks: fp32 torch encoder (the reference's own
    modules, restated) + CRF decode (C restatement) + left-pack.  Returns seconds per pass (best)."""
    from oracle import bonito_oracle as bo
    from oracle import cexact
    torch.set_num_threads(threads)
    if sd is None:
        sd = bo.reference_state_dict(n_base=N_BASE, seed=25)
    x = torch.randn(n_chunks, 1, CHUNK, generator=torch.Generator().manual_seed(1234))
    best = float('inf')
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            scores = bo.encoder_forward(sd, x, N_BASE, library=True)
        labels = cexact.crf_decode_threads(scores.numpy(), N_BASE, threads=threads)
        cexact.pack(labels, ALPHABET)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_chunks = 64                                     # = BASELINE.json configs[0]
    cpu_reference_pass(4, cores)                      # warm-up (thread pools, page-in)
    times = []
    for _ in range(args.warmup):
        cpu_reference_pass(n_chunks, cores)
    for _ in range(args.steps):
        times.append(cpu_reference_pass(n_chunks, cores))
    total = sum(times)
    value = n_chunks * CHUNK * len(times) / total
    line = {
        'impl': 'reference', 'metric': 'signal samples/sec basecalled (sup@v3.3 XNA)', 'value': value, 'unit': 'samples/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(workload_config(args.batch, CHUNK), sample='each step = %d chunks of the %d-chunk batch on the '
                                                                  'host CPU' % (n_chunks, args.batch)),
        'cpu_baseline': {'value': value, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d chunks x %d samples per step (torch fp32 encoder + C CRF decode, both on all '
                                   '%d host threads)' % (n_chunks, CHUNK, cores)},
        'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--dtype', default='fp16', choices=['fp16', 'bf16'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from oracle import bonito_oracle as bo            # weights generator + cpu_baseline leg only
    from xna_basecaller_b200._lib import Handle

    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    N, L, T = args.batch, CHUNK, T_STEPS
    h = Handle(ALPHABET, STATE_LEN, max_N=N, max_T=T, device=dev, bf16=(args.dtype == 'bf16'))
    sd = bo.reference_state_dict(n_base=N_BASE, seed=25)
    h.load_weights(sd)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(N, L, generator=g).pin_memory()
    x_dev = x_host.to(dev)
    seq_host = torch.empty(N, T, dtype=torch.int8).pin_memory()
    lens_host = torch.empty(N, dtype=torch.int32).pin_memory()
    scores = torch.empty(T, N, h.C * h.NZ, dtype=torch.float32, device=dev)

    def step_device():
        s = h.encoder(x_dev)
        return h.decode(s, want_qstring=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    h.set_profiling(True)
    h.stage_times()
    launches0 = h.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = h.launches - launches0
    stages = h.stage_times()
    h.set_profiling(False)

    # end to end through the host-buffer C-ABI call
    # (xb_compute_scores_submit / _wait: the pipelined form of xb_compute_scores_host that basecall() uses -- batch i+1 is
    # submitted before batch i is collected; every step still copies its input from pinned host memory and its result back)
    seq_hosts = [seq_host, torch.empty_like(seq_host).pin_memory()]
    lens_hosts = [lens_host, torch.empty_like(lens_host).pin_memory()]

    def run_e2e(k):
        for i in range(k):
            if i >= 2:
                h.compute_scores_wait(i & 1)
            h.compute_scores_submit(i & 1, x_host, seq_hosts[i & 1], lens_hosts[i & 1])
        for i in range(max(k - 2, 0), k):
            h.compute_scores_wait(i & 1)

    run_e2e(2)
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    clocks = sampler.summary()

    times = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = times.tolist()
    samples_per_step = N * L * world
    value = samples_per_step * args.steps / (dev_ms / 1e3)
    e2e_value = samples_per_step * args.steps / (e2e_ms / 1e3)

    if rank == 0:
        pk = peaks()
        rec_ms, rec_spans = stages['lstm_recurrence']
        # dominant kernel: the LSTM recurrence (one span = one layer = T recurrent steps of [N,768]x[768,3072])
        flops_per_span = T * 2.0 * N * FEATURES * 4 * FEATURES
        achieved = flops_per_span * rec_spans / (rec_ms / 1e3) / 1e12 if rec_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get('lstm_persistent_kernel', {}).get('dram_bytes_per_launch')

        def stage_ms(*names):
            return sum(stages[k][0] for k in names) / args.steps

        S_BYTES = 4 * h.C * h.NZ
        dec_ms = stage_ms('crf_alpha', 'crf_backward', 'crf_viterbi')
        inproj_ms = stage_ms('lstm_inproj_gemm')
        head_ms = stage_ms('crf_head_gemm')
        conv_ms = stage_ms('conv12_im2col', 'conv3_gemm')
        per_stage = {
            'lstm_inproj_gemm': {'bound': 'tensor', 'achieved': 5 * flops_per_span / (inproj_ms / 1e3) / 1e12,
                                 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s'},
            'crf_head_gemm': {'bound': 'hbm', 'achieved': (T * N * (h.C * h.NZ * 4 + FEATURES * 2)) / (head_ms / 1e3) / 1e9,
                              'peak': pk['hbm_gbs'], 'unit': 'GB/s'},
            'crf_decode': {'bound': 'hbm', 'achieved': (T * N * (2 * S_BYTES + 1)) / (dec_ms / 1e3) / 1e9,
                           'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                           'note': 'algorithmic bytes (2*S+1) per (t, chunk), S = %d B of fp32 scores' % S_BYTES},
            'conv_stem': {'bound': 'hbm', 'achieved': (N * (L * 4 + T * FEATURES * 2)) / (conv_ms / 1e3) / 1e9,
                          'peak': pk['hbm_gbs'], 'unit': 'GB/s'},
        }
        for v in per_stage.values():
            v['frac'] = v['achieved'] / v['peak']
        line = {
            'metric': 'signal samples/sec basecalled (sup@v3.3 XNA)', 'value': value, 'unit': 'samples/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': dev_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
            'config': workload_config(N, L),
            'e2e': {'value': e2e_value, 'unit': 'samples/s', 'h2d_bytes_per_step': N * L * 4 * world,
                    'd2h_bytes_per_step': (N * T + N * 4) * world},
            'gpu_launches': launches,
            'roofline': {'bound': 'tensor', 'kernel': 'lstm_recurrence', 'achieved': achieved, 'peak': pk['tf_sustained'],
                         'unit': 'TFLOP/s', 'frac': achieved / pk['tf_sustained'], 'traffic': traffic,
                         'peak_source': pk['source'] + ', sustained figure (kernel timed inside a long step)'},
            'stage_rooflines': per_stage,
            'stages_ms_per_step': {k: v[0] / args.steps for k, v in stages.items()},
            'model_tflops': (sum(FLOP_PER_CHUNK.values()) * N * world) / (dev_ms / args.steps / 1e3) / 1e12,
            'clocks': clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_chunks = 64                             # = BASELINE.json configs[0]
            cpu_reference_pass(4, cores)
            sec = cpu_reference_pass(n_chunks, cores, repeats=2)
            line['cpu_baseline'] = {'value': n_chunks * L / sec, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                                    'sample': '%d chunks x %d samples, best of 2 (torch fp32 encoder + C CRF decode, both on '
                                              'all %d host threads)' % (n_chunks, L, cores)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
