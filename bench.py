"""Headline benchmark: captions/sec at beam=5, seq_len=20 (BASELINE.json metric), UpDown, 36x2048 bottom-up features.

    python bench.py --gpus N --steps K --warmup W            # this engine (one process per GPU; torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A "step" = one pass of the hot path (AttModel._sample_beam: prologue + 20 timesteps + beam bookkeeping) over one batch of
synthetic inputs (configs[1]: batch 256 per GPU).  Images are independent, so ranks shard the work with no data-path
collective ("scaling": "weak"); the only collectives are the timing barrier and the max-over-ranks reduction.

  value   captions/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e     the same metric through the public model(...) call with HOST (pinned) inputs: H2D copy of the features and the
          D2H read of the caption ids are inside the timed region, every step
  roofline the dominant kernel (the persistent tcgen05 GEMM: CTA-pair kernel for the LSTM-gate and logit call sites): algorithmic
          FLOPs of all its launches / their CUDA-event time vs the measured bf16 tensor peak in MEASURED_PEAKS.json, the DRAM traffic
          of the largest call site from the committed ncu capture, and the fraction of the 3-pass ceiling
  cpu_baseline  the oracle port of the reference's CPU path, timed on this box's host cores on a bounded sample

Other workloads (--workload): transformer_beam / aoa_beam (BASELINE configs[2] shape and AoANet decode), updown_scst / aoa_scst (SCST
training step incl. H2D, the single NCCL gradient all-reduce and Adam; aoa_scst = BASELINE configs[3]).  The GPU arms build their
seeded random-init model and features from imagecaptioning.pytorch_b200.synthetic; only cpu_reference_rate() touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

CFG = dict(V=9487, E=1000, H=1000, A=512, F_fc=2048, F_att=2048, T=20)     # configs/updown/updown.yml + opts.py defaults
R = 36


def parse():
    p = argparse.ArgumentParser()
    p.add_argument('--gpus', type=int, default=1)
    p.add_argument('--steps', type=int, default=10)
    p.add_argument('--warmup', type=int, default=3)
    p.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    p.add_argument('--batch', type=int, default=None, help='images per GPU per step (default 256; 10 for updown_scst = BASELINE configs[3])')
    p.add_argument('--beam', type=int, default=5)
    p.add_argument('--mode', default='tc_f16x3', choices=['tc_f16x3', 'tc_f16x1', 'simt_fp32'])
    p.add_argument('--cpu-batch', type=int, default=32, help='images per CPU-baseline step (bounded sample)')
    p.add_argument('--no-cpu-baseline', action='store_true')
    p.add_argument('--workload', default='updown_beam', choices=['updown_beam', 'transformer_beam', 'aoa_beam', 'updown_scst', 'aoa_scst'],
                   help='updown_beam = BASELINE.json configs[1] (the headline); transformer_beam = configs[2] (use --batch 64); aoa_beam = AoANet decode')
    args = p.parse_args()
    if args.batch is None:
        args.batch = 10 if args.workload in ('updown_scst', 'aoa_scst') else 256
    return args


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        names = {'hw_slowdown': 0x8, 'sw_power_cap': 0x4, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20, 'hw_power_brake': 0x80}
        while not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                bits = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, 'nvmlDeviceGetCurrentClocksEventReasons') \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, b in names.items():
                    if bits & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(float(os.environ.get('CAPB200_CLOCK_SAMPLE_S', '0.05')))

    def summary(self):
        return {'sm_mhz': statistics.median(self.samples) if self.samples else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


def cpu_reference_rate(batch, beam, steps, warmup):
    """The oracle port (torch fp32) of AttModel._sample_beam on the same model / feature shapes.  The reference's eager loop
    of small GEMMs, sorts and gathers scales badly past a few dozen threads (0.7 captions/s with 128 threads vs ~25 with 8 on
    the same code), so the thread count is calibrated on a small batch and the best one is used and reported."""
    import torch
    from oracle import caption_oracle as co
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best = (None, float('inf'))
    Wc = co.make_weights('updown', CFG['V'], CFG['E'], CFG['H'], CFG['A'], CFG['F_fc'], CFG['F_att'], seed=1234, logit_scale=12.0)
    famc = co.Family('updown', Wc, 4)
    fcc, attc = co.make_inputs(8, R, CFG['F_fc'], CFG['F_att'], seed=1)
    for c in cands:
        torch.set_num_threads(c)
        with torch.no_grad():
            co.sample_beam(famc, fcc, attc, beam_size=beam)
            t0 = time.perf_counter()
            co.sample_beam(famc, fcc, attc, beam_size=beam)
            dt = time.perf_counter() - t0
        if dt < best[1]:
            best = (c, dt)
    torch.set_num_threads(best[0])
    W = co.make_weights('updown', CFG['V'], CFG['E'], CFG['H'], CFG['A'], CFG['F_fc'], CFG['F_att'], seed=1234, logit_scale=12.0)
    fam = co.Family('updown', W, CFG['T'])
    fc, att = co.make_inputs(batch, R, CFG['F_fc'], CFG['F_att'], seed=1234)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            co.sample_beam(fam, fc, att, beam_size=beam)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    dt = statistics.median(times)
    return batch / dt, dt, best[0]


def bench_scst(args, rank, world, local_rank, dev):
    """SCST samples/sec: UpDown, per-GPU batch = --batch images (BASELINE configs[3] uses 10), train_sample_n = 5, CIDEr-D reward, greedy
    baseline, BPTT, ONE gradient all-reduce (NCCL), Adam step.  value = 5 * batch * world / step time."""
    import argparse as ap
    import torch
    import torch.distributed as dist
    import imagecaptioning.pytorch_b200 as b200
    from imagecaptioning.pytorch_b200 import synthetic as syn
    B, n, T = args.batch, 5, CFG['T']
    aoa = args.workload == 'aoa_scst'
    if aoa:       # configs/aoa.yml: E = H = 1024, 8 heads, 6 refiner layers, ctx_drop, dropout_aoa 0.3 (BASELINE configs[3])
        model = syn.build_model('aoa', seed=1234, logit_scale=6.0, mode=args.mode, device=dev, heads=8, **dict(CFG, E=1024, H=1024, A=0))
    else:
        model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode=args.mode, device=dev, **CFG)
    fam_name = 'AoANet' if aoa else 'UpDown'
    model.train()
    df, ref_len = syn.document_frequency(syn.make_refs(1000, CFG['V'], seed=4))              # synthetic DF table (format of prepro_ngrams.py)
    b200.rewards.reset_scorer()
    b200.rewards.init_scorer(b200.rewards.CiderDTable(df, ref_len))
    opt = ap.Namespace(sc_sample_method='greedy', sc_beam_size=1, train_sample_method='sample', train_beam_size=1, train_sample_n=n,
                       cider_reward_weight=1, bleu_reward_weight=0)
    lw = b200.B200LossWrapper(model, opt)
    optim = torch.optim.Adam(model.parameters(), lr=5e-5)
    host = [syn.make_inputs(B, R, CFG['F_fc'], CFG['F_att'], seed=99 + 13 * rank + i) for i in range(3)]
    host = [(a.pin_memory(), b.pin_memory()) for a, b in host]
    gts = syn.make_refs(B, CFG['V'], seed=5 + rank)
    idx = torch.arange(B)
    grad_bytes = [0]

    def step(i):
        fc_h, att_h = host[i % 3]
        fc, att = fc_h.to(dev, non_blocking=True), att_h.to(dev, non_blocking=True)      # H2D every step (inputs start on the host)
        out = lw(fc, att, None, None, None, gts, idx, True, False, False)
        optim.zero_grad(set_to_none=True)
        out['loss'].backward()
        grad_bytes[0] = b200.parallel.allreduce_gradients(model.parameters())            # the one collective of the step
        torch.nn.utils.clip_grad_value_(model.parameters(), 0.1)
        optim.step()
        return float(out['loss'])                                                        # D2H read of the loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = model.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step_ms = []
    for i in range(args.steps):
        t_s = time.perf_counter()
        step(args.warmup + i)
        step_ms.append((time.perf_counter() - t_s) * 1e3)
    e1.record()
    barrier()
    if os.environ.get('CAPB200_BENCH_STEP_TIMES'):
        print('rank %d per-step wall ms: %s' % (rank, ' '.join('%.1f' % v for v in step_ms)), file=sys.stderr, flush=True)
    sampler.stop_flag = True
    sampler.join()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    if rank == 0:
        value = world * B * n * args.steps / (ms / 1e3)
        line = {'metric': 'SCST samples/sec (%s, train_sample_n=5, CIDEr-D reward, greedy baseline, BPTT, Adam)' % fam_name, 'value': value, 'unit': 'samples/s',
                'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': '%s SCST step, per-GPU batch=%d images x %d samples, 36x2048 feats, seq_len=20, V=9487' % (fam_name, B, n),
                           'images_per_sec': value / n, 'parallelism': 'dp%d, one gradient all-reduce of %d bytes per step' % (world, grad_bytes[0]),
                           'numeric_mode': 'greedy baseline %s (tcgen05); sampling + backward on 3xTF32 split-K tensor-core GEMMs over the fp32 weights, weight-gradient GEMMs fp32' % args.mode},
                'clocks': sampler.summary(),
                'e2e': {'value': value, 'unit': 'samples/s', 'h2d_bytes_per_step': B * (CFG['F_fc'] + R * CFG['F_att']) * 4, 'd2h_bytes_per_step': 4},
                'gpu_launches': model.launch_count - l0, 'roofline': None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    names = {'updown_beam': 'UpDown', 'transformer_beam': 'Transformer 6+6/512/2048/8', 'aoa_beam': 'AoANet 1024', 'updown_scst': 'UpDown SCST',
             'aoa_scst': 'AoANet SCST'}
    workload = '%s beam=%d, %dx2048 bottom-up feats, batch=%d per GPU, seq_len=20, V=9487' % (names[args.workload], args.beam, R, args.batch)

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, min(args.steps, 5))
        rate, dt, cores = cpu_reference_rate(args.cpu_batch, args.beam, steps, 1)
        line = {'impl': 'reference', 'metric': 'captions/sec at beam=5 seq_len=20', 'value': rate, 'unit': 'captions/s', 'n_gpus': args.gpus,
                'steps': steps, 'warmup': 1, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': {'workload': workload, 'sample': 'batch=%d per step on the host cores' % args.cpu_batch},
                'cpu_baseline': {'value': rate, 'unit': 'captions/s', 'cores': cores, 'kind': 'port',
                                 'host_cpus': os.cpu_count(),
                                 'sample': '%d steps of batch %d (oracle port of the reference CPU path, torch fp32, best thread count of a calibration sweep)' % (steps, args.cpu_batch)},
                'e2e': {'value': rate, 'unit': 'captions/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if local_rank == 0:
        ge.build()
    torch.cuda.set_device(local_rank)
    try:        # bind this rank to the CPU cores next to its GPU (NUMA): the SCST step is ~1600 launches of host-side work per step
        if os.environ.get('CAPB200_BENCH_NO_AFFINITY'):
            raise RuntimeError('disabled')
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist.barrier()
    from imagecaptioning.pytorch_b200 import synthetic as syn      # seeded random-init weights / features: the GPU arm never touches oracle/
    dev = torch.device('cuda', local_rank)
    if args.workload in ('updown_scst', 'aoa_scst'):
        return bench_scst(args, rank, world, local_rank, dev)
    if args.workload == 'updown_beam':
        model = syn.build_model('updown', seed=1234, logit_scale=12.0, mode=args.mode, device=dev, **CFG)
    elif args.workload == 'transformer_beam':     # configs/transformer/transformer.yml: d_model 512, d_ff 2048, 6 + 6 layers, 8 heads
        model = syn.build_model('transformer', seed=1234, logit_scale=3.0, mode=args.mode, device=dev, heads=8,
                                **dict(CFG, E=512, H=2048, A=6))
    else:                                         # configs/aoa.yml: E = H = 1024, 8 heads, 6 refiner layers
        model = syn.build_model('aoa', seed=1234, logit_scale=6.0, mode=args.mode, device=dev, heads=8, **dict(CFG, E=1024, H=1024, A=0))
    B, T = args.batch, CFG['T']
    opt = {'beam_size': args.beam, 'sample_n': 1}
    n_rot = 3                                         # rotate input batches; per-step working set (features, weights, 1 GB slab) >> 126 MB L2
    host = [syn.make_inputs(B, R, CFG['F_fc'], CFG['F_att'], seed=1234 + 17 * rank + i) for i in range(n_rot)]
    host = [(a.pin_memory(), b.pin_memory()) for a, b in host]
    devin = [(a.to(dev), b.to(dev)) for a, b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident(i):
        fc, att = devin[i % n_rot]
        with torch.no_grad():
            return model(fc, att, None, opt=opt, mode='sample')

    # End-to-end: features start in pinned HOST memory every step; the H2D copy of step i+1 is issued on a side stream while
    # step i decodes, and each step's caption ids are copied back to pinned host memory (D2H) and read one step later.
    copy_stream = torch.cuda.Stream(device=dev)
    pending = {}
    out_host = [torch.empty(B, T, dtype=torch.long).pin_memory() for _ in range(2)]
    out_events = [None, None]

    def prefetch(i):
        fc_h, att_h = host[i % n_rot]
        with torch.cuda.stream(copy_stream):
            fc = fc_h.to(dev, non_blocking=True)
            att = att_h.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending[i] = (fc, att, ev)

    def step_e2e(i):
        if i not in pending:
            prefetch(i)
        fc, att, ev = pending.pop(i)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        fc.record_stream(cur)
        att.record_stream(cur)
        prefetch(i + 1)
        with torch.no_grad():
            seq, _ = model(fc, att, None, opt=opt, mode='sample')
        slot = i % 2
        if out_events[slot] is not None:
            out_events[slot].synchronize()                 # the ids of step i-2 are on the host now
            _ = int(out_host[slot][0, 0])
        out_host[slot].copy_(seq, non_blocking=True)       # the captions (ids) are the step's result
        out_events[slot] = torch.cuda.Event()
        out_events[slot].record(cur)
        return seq

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        l0 = model.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        e1.record()
        barrier()
        sampler.stop_flag = True
        sampler.join()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), sampler.summary(), model.launch_count - l0

    ms, clocks, launches = timed(step_resident, args.steps, args.warmup)
    value = world * B * args.steps / (ms / 1e3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, max(1, args.warmup - 1))
    pending.clear()
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    if args.workload != 'updown_beam':
        if rank == 0:
            line = {'metric': 'captions/sec at beam=5 seq_len=20', 'value': value, 'unit': 'captions/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                    'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.mode, 'data': 'synthetic',
                    'config': {'workload': workload, 'numeric_mode': args.mode, 'global_batch': B * world}, 'clocks': clocks,
                    'e2e': {'value': e2e, 'unit': 'captions/s', 'h2d_bytes_per_step': B * (CFG['F_fc'] + R * CFG['F_att']) * 4, 'd2h_bytes_per_step': B * T * 8},
                    'gpu_launches': launches, 'roofline': None}
            print(json.dumps(line))
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel, timed live with CUDA events on the launching stream over a few more steps
    model.set_profiling(True)
    for i in range(3):
        step_resident(i)
    prof = model.read_profile()
    model.set_profiling(False)
    peaks_path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))['bf16_tflops_sustained']), 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)'
    else:
        peak, peak_src = 1400.0, 'fallback 1.4 PFLOP/s sustained (of fallback)'
    # The dominant kernel is the persistent tcgen05 GEMM (gemm_tc_kernel): every dense contraction of the step is a launch of it.
    # achieved = algorithmic FLOPs (2*M*N*K of the contraction actually executed) of ALL its launches / their summed CUDA-event time;
    # the largest single call site (language-LSTM gates, M=B*beam, N=4000, K=3000) is listed beside it.
    # DRAM bytes per launch of the largest call site (lang_lstm gates) from the committed `ncu --set full` capture, when present
    traffic, traffic_src = None, None
    tpath = os.path.join(REPO, 'profiles', 'roofline_traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic, traffic_src = tj.get('dram_bytes_per_launch'), tj.get('source')
    all_ms = sum(v[0] for v in prof.values())
    all_fl = sum(v[1] for v in prof.values())
    all_calls = sum(v[2] for v in prof.values())
    achieved = all_fl / (all_ms / 1e3) / 1e12 if all_ms > 0 else 0.0
    big_ms, big_fl, big_calls = prof['lang_lstm']
    roofline = {'bound': 'tensor', 'kernel': 'gemm_tc_pair_kernel<144,%d> / gemm_tc_kernel<64,..> (persistent tcgen05 GEMM, cta_group::2 pairs for the large call sites; all call sites of the step)' % (3 if args.mode == 'tc_f16x3' else 1),
                'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': peak_src,
                'mma_passes': 3 if args.mode == 'tc_f16x3' else 1,
                'frac_of_pass_ceiling': achieved / (peak / (3 if args.mode == 'tc_f16x3' else 1)),       # fp32-grade results cost 3 MMA passes per product
                'launches_timed': all_calls, 'avg_launch_ms': all_ms / max(all_calls, 1),
                'share_of_step': (all_ms / 3) / (ms / args.steps),
                'largest_call_site': {'name': 'lang_lstm gates M=%d N=4000 K=3000 (fused LSTM cell epilogue)' % (B * args.beam),
                                      'tflops': big_fl / (big_ms / 1e3) / 1e12 if big_ms > 0 else 0.0, 'avg_launch_ms': big_ms / max(big_calls, 1),
                                      'frac': (big_fl / (big_ms / 1e3) / 1e12 if big_ms > 0 else 0.0) / peak},
                'per_gemm_ms_per_step': {k: v[0] / 3 for k, v in prof.items() if v[2] > 0},
                'per_gemm_tflops': {k: v[1] / (v[0] / 1e3) / 1e12 for k, v in prof.items() if v[0] > 0}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {'metric': 'captions/sec at beam=5 seq_len=20', 'value': value, 'unit': 'captions/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (fp32-grade: split-fp16 x3 tensor-core passes, fp32 accumulate)' if args.mode == 'tc_f16x3' else args.mode, 'data': 'synthetic',
            'config': {'workload': workload, 'numeric_mode': args.mode, 'global_batch': B * world, 'parallelism': 'dp%d (independent images, no collective)' % world,
                       'l2': 'inputs rotated over %d batches; per-step working set ~1.3 GB >> 126 MB L2' % n_rot},
            'clocks': clocks,
            'e2e': {'value': e2e, 'unit': 'captions/s', 'h2d_bytes_per_step': B * (CFG['F_fc'] + R * CFG['F_att']) * 4, 'd2h_bytes_per_step': B * T * 8,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches, 'roofline': roofline}
    if not args.no_cpu_baseline and world == 1:
        rate, dt, cores = cpu_reference_rate(args.cpu_batch, args.beam, 2, 1)
        line['cpu_baseline'] = {'value': rate, 'unit': 'captions/s', 'cores': cores, 'kind': 'port', 'host_cpus': os.cpu_count(),
                                'sample': '2 steps of batch %d through the oracle port of the reference CPU path (torch fp32, best thread count of a calibration sweep)' % args.cpu_batch}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
