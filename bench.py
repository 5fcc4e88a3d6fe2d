#!/usr/bin/env python
"""Benchmark of the MILAN describe-neurons hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W]                 our arm (one rank per GPU under torchrun)
  python bench.py --impl reference [--steps K] [--warmup W]           reference arm: the CPU oracle port

A "step" = describing one batch of synthetic neurons (k = 15 exemplars of 3x224x224 + mask, beam = 50, LM/PMI
rerank, length 15): ResNet-101 pyramid encode -> attention-LSTM beam search -> LM rerank -> token ids.
  value : neurons/s with the uint8 exemplars already resident in HBM when the timed region starts: ONE
          `milan_describe_device` call over the neurons of all K steps (decode of chunk i under the encoder of i+1)
  e2e   : the same through `milan_describe_host` with HOST (pinned) buffers: ONE call describes the neurons of all
          K steps (the call a user makes for a whole exemplar set); the H2D copy of every step's exemplars and the
          D2H read of the token ids / scores are inside the timed region (the engine overlaps the copy of chunk
          i+1 with the compute of chunk i)
Weak scaling: every rank describes its own shard of neurons; the only collective is an all-gather of the final
token ids + scores (NCCL), inside the timed region.
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

import torch  # noqa: E402

METRIC = 'neurons described/sec (k=15, beam=50)'
UNIT = 'neurons/s'
K_EXEMPLARS, BEAM, LENGTH, GROUP = 15, 50, 15, 16
MAX_STEPS_PER_CALL = 8
CONV_FLOP_PER_NEURON = 2.0 * 7.79935744e9 * K_EXEMPLARS  # 104 convs through layer4
# every conv input / residual / output tensor touched once as (hi, lo) bf16 planes; the four
# downsample tensors are never materialised (conv3 + downsample run as one GEMM): 157.5 - 11.6 + 2.1 GB per 64 neurons
CONV_BYTES_PER_NEURON = 148.0e9 / 64
SURVEY_BYTES_PER_NEURON = 0.97e9  # every conv tensor once as ONE bf16 NHWC plane
WORKLOAD = ('alexnet/imagenet 1k neurons, k=15, beam=50 + PMI rerank (synthetic exemplars of that shape, '
            'random-init MILAN resnet101 encoder + attention-LSTM decoder + LSTM LM, V=5004)')


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as handle:
            peaks = json.load(handle)
        return peaks, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '200'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        clocks, maxes, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in self.lines:
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 9:
                continue
            try:
                clocks.append(float(parts[1]))
                maxes.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        if not clocks:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        clocks.sort()
        return {'sm_mhz': clocks[len(clocks) // 2], 'sm_max_mhz': max(maxes), 'reasons': sorted(reasons),
                'samples': len(clocks)}


def shared_config():
    """The `config` both arms print (identical dicts: the driver compares them); arm-specific detail goes in `arm`."""
    return {'workload': WORKLOAD, 'k': K_EXEMPLARS, 'beam': BEAM, 'length': LENGTH, 'strategy': 'rerank',
            'temperature': 0.2, 'decode_batch': GROUP}


def bench_exemplars(nb, rank, index):
    """Synthetic uint8 exemplars of bench batch `index` of `rank` — the SAME bytes for our arm, the reference arm and
    the in-run parity check."""
    from neuron_descriptions_b200 import synthetic
    return synthetic.synthetic_exemplars(nb, K_EXEMPLARS, seed=1000 * rank + index)


def cpu_oracle_describe(images_u8, masks_u8, threads):
    """The reference's CPU path (oracle port, torch fp32) on the given exemplars: (tokens, scores, seconds)."""
    from neuron_descriptions_b200 import synthetic
    from oracle import milan_oracle as O
    torch.set_num_threads(threads)
    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0)
    images, masks = O.to_float_inputs(images_u8, masks_u8)
    tokens, scores = [], []
    start = time.perf_counter()
    with torch.no_grad():
        for lo in range(0, len(images), GROUP):  # Decoder.predict's batches, src/milan/decoders.py:809-871
            feats = O.encode(images[lo:lo + GROUP], masks[lo:lo + GROUP], sd)
            out = O.decode(feats, sd, vocab, strategy='rerank', beam_size=BEAM, length=LENGTH, temperature=0.2)
            pad = out.tokens.new_full((len(out.tokens), LENGTH), len(vocab) + 1)  # early exit: <stop> padding
            pad[:, :out.tokens.shape[1]] = out.tokens
            tokens.append(pad)
            scores.append(out.scores)
    elapsed = time.perf_counter() - start
    return torch.cat(tokens), torch.cat(scores), elapsed


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = GROUP  # one reference batch (16 neurons, ~6 s of CPU work) per step
    images_u8, masks_u8 = bench_exemplars(4 * GROUP, 0, 0)  # our arm's batch 0 of rank 0
    warm_sample = 2
    for _ in range(args.warmup):
        cpu_oracle_describe(images_u8[:warm_sample], masks_u8[:warm_sample], threads)
    times = []
    for i in range(args.steps):
        lo = (i % 4) * GROUP
        _, _, elapsed = cpu_oracle_describe(images_u8[lo:lo + sample], masks_u8[lo:lo + sample], threads)
        times.append(elapsed)
    total = sum(times)
    value = sample * len(times) / total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times), 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': shared_config(),
        'arm': {'neurons_per_step': sample, 'device': 'cpu', 'threads': threads,
                'warmup_sample': f'{warm_sample} neurons per warm-up step'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                         'sample': f'{sample} neurons per step x {len(times)} steps, oracle port of the reference '
                                   f'PyTorch path (torch fp32, {threads} threads), batch 16, rerank beam 50'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


CONFIG3_NEURONS = 64 + 256 + 512 + 1024 + 2048  # resnet152/places365: conv1 + layer1..4 (src/exemplars/models.py:321-326)


class TiledExemplars:
    """`CONFIG3_NEURONS` exemplar sets in the shape `Decoder.predict`'s uint8 fast path reads (`batch_u8`,
    `alloc_batch_u8`, `k`): neuron i is distinct set i % len(images), so that 11.8 GB of synthetic bytes need not be
    generated and held (the engine reads every byte it is given either way)."""
    transform_images = transform_masks = None

    def __init__(self, images_u8, masks_u8, n):
        self.images, self.masks, self.n, self.k = images_u8, masks_u8, n, images_u8.shape[1]

    def __len__(self):
        return self.n

    def alloc_batch_u8(self, n):
        return (torch.empty((n, *self.images.shape[1:]), dtype=torch.uint8, pin_memory=True),
                torch.empty((n, *self.masks.shape[1:]), dtype=torch.uint8, pin_memory=True))

    def batch_u8(self, lo, hi, out=None):
        out = out if out is not None else self.alloc_batch_u8(hi - lo)
        index = torch.arange(lo, hi) % len(self.images)
        torch.index_select(self.images, 0, index, out=out[0][:hi - lo])
        torch.index_select(self.masks, 0, index, out=out[1][:hi - lo])
        return out[0][:hi - lo], out[1][:hi - lo]


def strong_scaling_config3(sd, vocab, host, device, world, rank, precision):
    """BASELINE config 3 as STRONG scaling: 3904 neurons in total, sharded over the ranks, through the Python facade
    (`Decoder.predict` via `sharding.predict_sharded`: host uint8 feed, H2D, beam + rerank, all-gather, detokenise).
    The same path `scripts/compute_milan_descriptions.py` drives (scripts/config3_scaling.py measures it through the
    CLI on an on-disk set); here it runs inside the bench so that its per-N numbers reach the driver's records."""
    import torch.distributed as dist
    from neuron_descriptions_b200 import milan, sharding
    from neuron_descriptions_b200.milan import lang
    indexer = lang.Indexer(lang.Vocab(vocab), start=True, stop=True, pad=True, unk=True)
    decoder = milan.Decoder(indexer, milan.PyramidConvEncoder('resnet101', pretrained=False),
                            lm=milan.LanguageModel(indexer), precision=precision)
    decoder.load_state_dict(sd)
    decoder.to(device)
    dataset = TiledExemplars(torch.cat([host[0][0], host[1][0]]), torch.cat([host[0][1], host[1][1]]), CONFIG3_NEURONS)
    times = []
    for _ in range(2):  # the first pass includes one-off allocations, like a fresh CLI process does
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        start = time.perf_counter()
        captions = sharding.predict_sharded(decoder, dataset, world=world, rank=rank, strategy='rerank', beam_size=BEAM,
                                            temperature=0.2, device=device)
        torch.cuda.synchronize(device)
        elapsed = torch.tensor([time.perf_counter() - start], device=device)
        if world > 1:
            dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
        times.append(float(elapsed.item()))
    assert len(captions) == CONFIG3_NEURONS
    decoder.engine.close()
    return {'workload': 'resnet152/places365 4k neurons (conv1 + layer1-4 = 3904 units), k=15, beam=50 + PMI rerank, '
                        'neuron-sharded: total work fixed as GPUs are added',
            'scaling': 'strong', 'neurons': CONFIG3_NEURONS, 'n_gpus': world, 'describe_s': times[1],
            'value': CONFIG3_NEURONS / times[1], 'unit': UNIT, 'neurons_per_hour': 3600.0 * CONFIG3_NEURONS / times[1],
            'first_pass_s': times[0],
            'path': 'Decoder.predict through sharding.predict_sharded (host uint8 feed, H2D, all-gather of token ids, '
                    'detokenisation), wall clock, max over ranks'}


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=6)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='ours', choices=('ours', 'reference'))
    parser.add_argument('--neurons-per-step', type=int, default=64)
    parser.add_argument('--precision', default='split', choices=('split', 'fast'))
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--no-fast-mode', action='store_true', help='skip the informational plain-bf16 measurement')
    parser.add_argument('--no-strong-scaling', action='store_true', help='skip the config-3 (3904 neurons) strong-scaling pass')
    args = parser.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import torch.distributed as dist
    from neuron_descriptions_b200 import _lib, synthetic
    from neuron_descriptions_b200.engine import Engine

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the milan_b200 engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        if os.environ.get('NCCL_DEBUG', '').upper() in ('VERSION', 'WARN'):
            os.environ.pop('NCCL_DEBUG')  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group('nccl', device_id=device)
    steps, warmup = args.steps, max(args.warmup, 3)
    nb = args.neurons_per_step

    vocab = synthetic.synthetic_vocab(5000)
    sd = synthetic.synthetic_state_dict(seed=0, sharpen=12.0, stop_bias=0.0)
    engine = Engine(sd, vocab_size=len(vocab) + 4, device=device, precision=args.precision, max_neurons=nb)
    lib = _lib.load()

    # Two distinct host batches per rank, alternated, so consecutive steps never see the same bytes; each step moves
    # nb*15*(3+1)*224*224 B (193 MB at nb=64) of inputs and GBs of activations: far beyond the 126 MB L2.
    host = []
    for i in range(2):
        images_u8, masks_u8 = bench_exemplars(nb, rank, i)
        host.append((images_u8.pin_memory(), masks_u8.pin_memory()))
    dev = [(im.to(device), mk.to(device)) for im, mk in host]
    gathered_tokens = torch.empty(world, nb, LENGTH, dtype=torch.long, device=device) if world > 1 else None
    gathered_scores = torch.empty(world, nb, dtype=torch.float32, device=device) if world > 1 else None

    def step_resident(i, engine_=None):  # one step as two calls (used for the informational fast-mode run)
        images, masks = dev[i % 2]
        eng = engine_ or engine
        feats = eng.encode(images.view(-1, 3, 224, 224), masks.view(-1, 1, 224, 224)).view(nb, K_EXEMPLARS, -1)
        _, _, _, tokens, scores, _ = eng.decode_beam(feats, LENGTH, BEAM, True, 0.2, group_size=GROUP)
        if world > 1:
            dist.all_gather_into_tensor(gathered_tokens, tokens)
            dist.all_gather_into_tensor(gathered_scores, scores)
        return tokens

    # value / e2e: the exemplars of the K steps (alternating the two batches) go through the pipelined API in calls
    # of at most MAX_STEPS_PER_CALL steps, so a large --steps does not need K x 193 MB of pinned / device buffers.
    per_call = min(steps, MAX_STEPS_PER_CALL)
    call_steps = [min(per_call, steps - lo) for lo in range(0, steps, per_call)]
    dev_all = (torch.cat([dev[i % 2][0] for i in range(per_call)]), torch.cat([dev[i % 2][1] for i in range(per_call)]))
    host_all = (torch.cat([host[i % 2][0] for i in range(per_call)]).pin_memory(),
                torch.cat([host[i % 2][1] for i in range(per_call)]).pin_memory())
    gathered_all_tokens = torch.empty(world, nb * per_call, LENGTH, dtype=torch.long, device=device) if world > 1 else None
    gathered_all_scores = torch.empty(world, nb * per_call, dtype=torch.float32, device=device) if world > 1 else None

    def gather(tokens, scores):
        if world > 1:
            n = tokens.shape[0]
            if n == nb * per_call:
                dist.all_gather_into_tensor(gathered_all_tokens, tokens.to(device, non_blocking=True))
                dist.all_gather_into_tensor(gathered_all_scores, scores.to(device, non_blocking=True))
            else:  # the short last call of an uneven split
                dist.all_gather_into_tensor(gathered_all_tokens[:, :n].contiguous(), tokens.to(device))
                dist.all_gather_into_tensor(gathered_all_scores[:, :n].contiguous(), scores.to(device))

    def resident_all(engine_):
        for n_steps in call_steps:
            n = nb * n_steps
            tokens, scores = engine_.describe_device(dev_all[0][:n], dev_all[1][:n], strategy='rerank', length=LENGTH,
                                                     beam=BEAM, group_size=GROUP, temperature=0.2)
            gather(tokens, scores)
        return tokens

    e2e_first = {}

    def e2e_all(engine_):
        for n_steps in call_steps:
            n = nb * n_steps
            tokens, scores, _ = engine_.describe_host(host_all[0][:n], host_all[1][:n], strategy='rerank', length=LENGTH,
                                                      beam=BEAM, group_size=GROUP, temperature=0.2)
            gather(tokens, scores)
            e2e_first.setdefault('out', (tokens, scores))  # kept for the in-run parity check (first call, batch 0 first)
        return tokens

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(fn, profile, single_call=False):
        """Time K steps: K calls of fn(i), or (single_call) one call of fn() that covers all K steps."""
        if single_call:
            fn()  # warm-up: the same K (>= 3) steps once
        else:
            for i in range(warmup):
                fn(i)
        barrier()
        engine.set_profiling(profile)
        launches0 = lib.milan_launch_count()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local_rank) if rank == 0 else None
        if sampler:
            sampler.start()
        start.record()
        if single_call:
            fn()
        else:
            for i in range(steps):
                fn(warmup + i)
        end.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = start.elapsed_time(end)
        launches = lib.milan_launch_count() - launches0
        prof = engine.get_profile() if profile else None
        engine.set_profiling(False)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, prof, clocks

    # value: pipelined call, no per-kernel events. Roofline: a second pass over the same K steps with the engine's
    # profiling on, which records CUDA events around every conv launch and runs encode / decode back to back (the
    # overlapped decode kernels would otherwise stretch the event-timed conv durations).
    ms_res, launches, _, clocks = timed(lambda: resident_all(engine), False, single_call=True)
    ms_prof, _, prof, clocks_prof = timed(lambda: resident_all(engine), True, single_call=True)
    ms_e2e, _, _, clocks_e2e = timed(lambda: e2e_all(engine), False, single_call=True)

    fast = None
    if args.precision == 'split' and not args.no_fast_mode:
        # Informational only: the same step with plain bf16 operands (1 MMA per k-block). NOT parity-green
        # (profiles/*parity_report*: log-prob errors up to O(1)), so it is never the headline value.
        parity_engine = engine
        engine = Engine(sd, vocab_size=len(vocab) + 4, device=device, precision='fast', max_neurons=nb)
        ms_fast, _, prof_fast, _ = timed(lambda: resident_all(engine), True, single_call=True)
        fast = {'value': nb * steps * world / (ms_fast / 1e3), 'unit': UNIT, 'ms_per_step': ms_fast / steps,
                'encoder_convs_ms_per_step': prof_fast['conv_ms'] / steps,
                'roofline_frac': CONV_FLOP_PER_NEURON * nb * steps / (prof_fast['conv_ms'] / 1e3) / 1e12 /
                measured_peaks()[0]['bf16_tflops_sustained'],
                'note': 'plain bf16 operands; fails the 1e-3 parity bar, reported for context only'}
        engine.close()
        engine = parity_engine

    total_neurons = nb * steps * world
    value = total_neurons / (ms_res / 1e3)
    e2e_value = total_neurons / (ms_e2e / 1e3)
    peaks, peak_kind = measured_peaks()
    conv_ms = prof['conv_ms']
    conv_launches = max(1, prof['conv_launches'])
    achieved = CONV_FLOP_PER_NEURON * nb * steps / (conv_ms / 1e3) / 1e12  # algorithmic TFLOP/s while convs run
    peak = peaks['bf16_tflops_sustained']
    traffic_path = os.path.join(ROOT, 'profiles', 'conv_traffic.json')
    traffic = None
    if os.path.exists(traffic_path):
        with open(traffic_path) as handle:
            traffic = json.load(handle).get('dram_bytes_per_launch')

    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms_res / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x3' if args.precision == 'split' else 'bf16', 'data': 'synthetic',
        'config': shared_config(),
        'arm': {
            'neurons_per_step_per_gpu': nb, 'parallelism': f'neuron-sharded x{world}',
            'precision': ('bf16 hi/lo split operands (3 products per K step, issued as 2 or 3 tcgen05 MMAs), fp32 TMEM '
                          'accumulation (fp32-class results; parity-tested)' if args.precision == 'split'
                          else 'plain bf16 operands'),
            'conv_cta_pairs': os.environ.get('MILAN_PAIR', '1') not in ('', '0'),  # cta_group::2 conv kernel (default on)
            'l2': 'inputs larger than L2: 193 MB of fresh exemplars + >10 GB of activations per step vs 126 MB L2',
            'call': f'milan_describe_device over the {steps} steps ({nb * steps} resident neurons) in {len(call_steps)} call(s)',
        },
        'roofline': {
            'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
            'traffic': traffic, 'kernel': 'conv_gemm_kernel (the 104 encoder convolutions of a step, 100 launches)',
            'peak_kind': f'{peak_kind} bf16 sustained (kernel timed inside a long step)',
            'algorithmic_flop_per_launch': CONV_FLOP_PER_NEURON * nb * steps / conv_launches,
            'avg_launch_ms': conv_ms / conv_launches, 'conv_share_of_step': conv_ms / ms_prof,
            'timed_in': f'a second pass over the same {steps} steps with per-launch CUDA events and the decode overlap '
                        f'off ({ms_prof / steps:.2f} ms per step, SM clock {clocks_prof["sm_mhz"] if clocks_prof else None} MHz)',
            'mma_flop_multiplier': 3 if args.precision == 'split' else 1,
            'tensor_pipe_frac_incl_split': achieved * (3 if args.precision == 'split' else 1) / peak,
            # bytes THIS implementation has to move (activations as hi + lo bf16 planes = fp32 bytes, each touched once)
            'implementation_bytes_per_launch': CONV_BYTES_PER_NEURON * nb * steps / conv_launches,
            # 0.97 GB per neuron with single-plane bf16 NHWC activations
            'survey_algorithmic_bytes_per_launch': SURVEY_BYTES_PER_NEURON * nb * steps / conv_launches,
            'hbm_gbs_while_convs_run': CONV_BYTES_PER_NEURON * nb * steps / (conv_ms / 1e3) / 1e9,
            'hbm_peak_gbs': peaks['hbm_gbs'],
        },
        'phases_ms_per_step': {'encoder_convs': conv_ms / steps, 'step_total': ms_res / steps,
                               'step_total_profiled_pass': ms_prof / steps},
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': nb * K_EXEMPLARS * 4 * 224 * 224,
                'd2h_bytes_per_step': nb * (LENGTH * 8 + 4 + 4), 'ms_per_step': ms_e2e / steps,
                'call': f'milan_describe_host over the {steps} steps ({nb * steps} neurons, pinned host buffers) in '
                        f'{len(call_steps)} call(s); H2D of chunk i+1 overlaps the compute of chunk i'},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'clocks_e2e': clocks_e2e,
    }
    if fast is not None:
        line['fast_mode'] = fast
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # cpu_baseline AND in-run parity check: the oracle describes the first two reference batches of the timed
        # run's own exemplars (batch 0 of this rank); its token ids / scores are compared with what the timed
        # `milan_describe_host` call returned for those neurons.
        threads = os.cpu_count() or 1
        sample = min(2 * GROUP, nb)  # ~10-15 s of CPU work
        ref_tokens, ref_scores, elapsed = cpu_oracle_describe(host[0][0][:sample], host[0][1][:sample], threads)
        line['cpu_baseline'] = {'value': sample / elapsed, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                'sample': f'the first {sample} neurons ({sample * K_EXEMPLARS} exemplars) of the timed run through '
                                          f'the oracle port of the reference PyTorch CPU path, rerank beam 50, {elapsed:.1f} s'}
        got_tokens, got_scores = (t.cpu() for t in e2e_first['out'])
        got_tokens, got_scores = got_tokens[:sample], got_scores[:sample]
        same = (got_tokens == ref_tokens).all(dim=1)
        err = (got_scores - ref_scores).abs()
        # identical ids, or (near-tie in the rerank argmax) a different sequence whose score is within tolerance
        ok = bool((err <= 1e-3).all())
        line['parity_check'] = {'neurons': sample, 'sequences_identical': int(same.sum()),
                                'max_abs_score_err': float(err.max()), 'tolerance': 1e-3, 'ok': ok,
                                'against': 'oracle port on the same exemplars, same run'}
        if not ok:
            raise SystemExit(f'bench.py: timed outputs differ from the oracle: {line["parity_check"]}')
    engine.close()
    if args.precision == 'split' and not args.no_strong_scaling:
        line['strong_scaling_config3'] = strong_scaling_config3(sd, vocab, host, device, world, rank, args.precision)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
