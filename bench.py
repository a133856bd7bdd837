#!/usr/bin/env python
"""Benchmark of the UNet -> WS hot path (BASELINE.json metric: UNet-WS 512x512 images/sec; conv tensor-pipe
fraction; estimator HBM fraction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input: PER_GPU (default 256) 512x512 uint8 LSBr
(alpha sweep) stego images per GPU through the fused UNet(unet_2, random init) -> WS beta_hat chain
(BASELINE.json configs[2]; at N > 1 every rank processes its own shard = configs[3], weak scaling, one NCCL
all_gather of beta_hat per step). `value` is images/s with inputs resident in HBM; `e2e` is the same metric through
the host-buffer C-ABI call (pinned host images -> H2D -> chain -> D2H of beta_hat/l1 inside the timed region).
`--impl reference` times the reference's CPU implementation of the same per-image path (oracle/torch_port.py: the
same torch ATen CPU calls the reference makes, batch 1, autograd on) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMG_512 = 202.199          # algorithmic 2*MACs of unet_2 at 512x512 (exact)
EST_BYTES_PER_IMG = 512 * 512 + 4    # uint8 image in, float beta_hat out
ALPHAS = [0.01, 0.05, 0.1, 0.2, 0.4, 1.0]
METRIC = 'UNet-WS 512x512 images/sec'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tf_burst': p['bf16_tflops'], 'tf_sustained': p['bf16_tflops_sustained'], 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except (ValueError, IndexError):
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def make_inputs(n, device, size=512):
    from ws_unet_b200 import data as wdata
    per = (n + len(ALPHAS) - 1) // len(ALPHAS)
    parts = [wdata.synthetic_stego_fast(per, a, size, size, device, seed=i, unique=32 if size <= 512 else 8) for i, a in enumerate(ALPHAS)]
    import torch
    return torch.cat(parts)[:n].contiguous()


def build_model(device, seed=1234):
    import torch
    import ws_unet_b200 as W
    torch.manual_seed(seed)
    return W.get_model('unet_2', in_channels=1, out_channels=1, channel=[0], drop_rate=0.).to(device)


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline_run(seconds=15.0, min_images=3, no_grad=False):
    """The reference's per-image path on the host cores: batch 1, FP32, autograd on, all torch threads."""
    import numpy as np
    import torch
    from oracle import torch_port
    from ws_unet_b200 import data as wdata
    import ws_unet_b200 as W
    torch.manual_seed(1234)
    model = W.get_model('unet_2', 1, 1, [0], 0.)
    imgs = [wdata.embed_lsbr(wdata.synthetic_cover(i), 0.4, i).numpy() for i in range(2)]
    torch_port.predict_unet(imgs[0], model, no_grad=no_grad)  # warm-up (first call pays oneDNN primitive creation)
    t0 = time.perf_counter()
    n = 0
    while n < min_images or time.perf_counter() - t0 < seconds:
        torch_port.predict_unet(imgs[n % 2], model, no_grad=no_grad)
        n += 1
        if n >= 64:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    per_step = 2
    from oracle import torch_port
    from ws_unet_b200 import data as wdata
    import ws_unet_b200 as W
    torch.manual_seed(1234)
    model = W.get_model('unet_2', 1, 1, [0], 0.)
    imgs = [wdata.embed_lsbr(wdata.synthetic_cover(i), ALPHAS[i % 6], i).numpy() for i in range(per_step)]
    for _ in range(args.warmup):
        torch_port.predict_unet(imgs[0], model)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for im in imgs:
            torch_port.predict_unet(im, model)
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'UNet-WS (unet_2 random init) on 512x512 LSBr alpha-sweep stego images, reference CPU path: '
                               'batch 1, FP32, autograd on', 'images_per_step': per_step},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps} steps x {per_step} images, torch ATen CPU ops as called by src/unet/evaluate.py:31-52,125-132'},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ws_unet_b200 as W
    from ws_unet_b200 import _native, parallel

    # Libraries write banners to stdout (NCCL prints its version line there): keep the real stdout for the one JSON line
    # and point fd 1 at stderr for everything else.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = parallel.init_from_env()
    if world != args.gpus:
        raise SystemExit(f'--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}')
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    lib = _native.load()
    pk = peaks()
    per_gpu = args.per_gpu
    model = build_model(dev)
    S = args.size
    imgs = make_inputs(per_gpu, dev, S)
    if args.micro_batch:
        model.set_micro_batch(args.micro_batch, dev)
    n_total = per_gpu * world

    # measured first, as its own workload, before the tensor-core chain heats the part up
    est = None
    if rank == 0 and args.est_images > 0:
        # ---- estimator (HBM-bound, BASELINE.json configs[1]): KB-filter WS on resident uint8 images
        n_est = args.est_images
        est_sampler = ClockSampler(local)
        est_sampler.start()
        est_imgs = imgs.repeat((n_est + per_gpu - 1) // per_gpu, 1, 1, 1)[:n_est].contiguous()
        est = {}
        # nvidia-smi takes ~1 s to come up and holds driver locks while it does: launches stall behind it, which a
        # 0.4 ms kernel shows. Warm up until its first sample has arrived, then time.
        t_wait = time.perf_counter()
        while not est_sampler.rows and time.perf_counter() - t_wait < 5.0:
            W.ws_estimate(est_imgs, 'KB', weighted=0)
            torch.cuda.synchronize()
        for weighted in (0, 1):
            for _ in range(10):   # also lets the clocks ramp before the timed repetitions
                W.ws_estimate(est_imgs, 'KB', weighted=weighted)
            torch.cuda.synchronize()
            groups, reps = [], 100   # long groups: the host runs ahead of the GPU and absorbs nvidia-smi polling stalls
            for _ in range(3):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(reps):
                    W.ws_estimate(est_imgs, 'KB', weighted=weighted)
                a1.record()
                torch.cuda.synchronize()
                groups.append(a0.elapsed_time(a1) / reps / 1e3)
            sec = sorted(groups)[1]   # median of three groups of 100 back-to-back calls
            gbs = (S * S + 4) * n_est / sec / 1e9
            est[f'kb_w{weighted}'] = {'images_per_s': n_est / sec, 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                                       'frac': gbs / pk['hbm_gbs'], 'bound': 'hbm', 'images': n_est,
                                       'kernel': 'filter_ws_adjoint_kernel + finalize_kernel' if weighted == 0
                                       else 'filter_ws_window_kernel + finalize_kernel'}
        # the same estimator end to end from pinned host memory (H2D copy inside the timed region): PCIe-bound
        n_h = min(n_est, 2048)
        host_est = est_imgs[:n_h].cpu().pin_memory()
        for _ in range(2):
            W.ws_estimate_host(host_est, 'KB', weighted=0)
        t0 = time.perf_counter()
        for _ in range(3):
            W.ws_estimate_host(host_est, 'KB', weighted=0)
        dt = (time.perf_counter() - t0) / 3
        est['kb_w0_e2e'] = {'images_per_s': n_h / dt, 'h2d_gbs': S * S * n_h / dt / 1e9, 'images': n_h,
                            'note': 'wsu_filter_ws_estimate_host: pinned host uint8 -> H2D -> kernel -> D2H, bound by the host link'}
        del est_imgs, host_est
        est['clocks'] = est_sampler.stop()
        torch.cuda.empty_cache()

    def step():
        beta = W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
        return parallel.gather_shards(beta, n_total) if world > 1 else beta

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.wsu_launch_count(1)
    h = model.native_handle(dev)
    # per-layer CUDA events (on the launching stream) stay on during the timed steps: the roofline below is taken from
    # the last micro-batch INSIDE the timed region, at the clocks the step sustains (24 event records per micro-batch)
    lib.wsu_set_option(h, b'profile', 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    launches = lib.wsu_launch_count(0)
    prof_buf = (ctypes.c_float * 64)()
    n_layers = lib.wsu_profile_read(h, prof_buf, 64) if rank == 0 else 0
    prof_info = ctypes.c_int64()
    lib.wsu_get_info(h, b'last_images', ctypes.byref(prof_info))
    last_mb = int(prof_info.value)
    lib.wsu_set_option(h, b'profile', 0)
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = n_total * args.steps / (ms_max / 1e3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory -> H2D -> chain -> D2H)
    host_img = imgs.cpu().pin_memory()
    host_out = torch.empty(2, per_gpu, dtype=torch.float32).pin_memory()

    def e2e_step():
        _native.check(lib.wsu_unet_ws_estimate_host(h, ctypes.c_void_p(host_img.data_ptr()), per_gpu, S, S, 0, 1, 1,
                                                    ctypes.c_void_p(host_out[0].data_ptr()), ctypes.c_void_p(host_out[1].data_ptr())))

    for _ in range(max(1, args.warmup // 2)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_total * args.steps / t.item()
    e2e_ok = bool(torch.allclose(host_out[0].to(dev), out[rank * per_gpu:(rank + 1) * per_gpu] if world > 1 else out, atol=0, rtol=0))

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- per-layer device times of the last micro-batch of the timed region -> roofline of the tensor-core chain
    buf = prof_buf
    info = ctypes.c_int64()
    lib.wsu_get_info(h, b'micro_batch', ctypes.byref(info))
    mb = int(info.value)
    layer_gflop = {'e11': 0.302, 'e12': 19.327, 'e21': 9.664, 'e22': 19.327, 'e31': 9.664, 'e32': 19.327, 'upconv3': 4.295,
                   'd31': 38.655, 'd32': 19.327, 'upconv4': 4.295, 'd41': 38.655, 'd42': 19.327 + 0.034}
    layers = []
    for i in range(max(0, n_layers)):
        name = lib.wsu_profile_name(h, i).decode()
        tf = layer_gflop.get(name, 0.0) * (S / 512) ** 2 * last_mb / (buf[i] * 1e-3) / 1e3 if buf[i] > 0 else 0.0
        layers.append({'layer': name, 'ms': round(buf[i], 4), 'tflops': round(tf, 1)})
    conv_ms = sum(l['ms'] for l in layers if l['layer'] != 'e11')
    conv_gflop = sum(layer_gflop[l['layer']] for l in layers if l['layer'] != 'e11') * (S / 512) ** 2
    achieved = conv_gflop * last_mb / (conv_ms * 1e-3) / 1e3 if conv_ms else 0.0

    # ---- CPU baseline (reference's per-image path on this box's host cores), bounded sample; N = 1 only: under torchrun
    # every rank is pinned to one OpenMP thread, which would not be the reference's configuration
    cpu_base = None
    if world == 1:
        cpu_v, cpu_n, cores = cpu_baseline_run(seconds=args.cpu_seconds)
        cpu_base = {'value': cpu_v, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                    'sample': f'{cpu_n} images 512x512, batch 1, FP32, autograd on, torch ATen CPU ops exactly as '
                              'src/unet/evaluate.py:31-52,125-132 calls them (oracle/torch_port.py)'}

    line = {
        'metric': METRIC if S == 512 else f'UNet-WS {S}x{S} images/sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16x3', 'data': 'synthetic',
        'config': {'workload': f'UNet-WS (unet_2 random init) beta_hat on {per_gpu} synthetic {S}x{S} uint8 LSBr stego images per '
                               f'GPU, alpha sweep {ALPHAS}, weighted=0 (BASELINE configs[2]; sharded by image at N>1 = configs[3])',
                   'images_per_step': n_total, 'micro_batch': mb, 'parallelism': f'image-sharded x{world}, all_gather(beta_hat)',
                   'arithmetic': 'split-bf16 operands (hi + lo), 3 tcgen05 MMAs per MAC, fp32 accumulation in TMEM; integer WS arithmetic',
                   'l2_policy': 'working set per micro-batch (>= 8 GB of activations) far exceeds the 126 MB L2; no flush needed'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': per_gpu * S * S * world,
                'd2h_bytes_per_step': per_gpu * 8 * world, 'matches_device_path': e2e_ok},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                     'frac': achieved / pk['tf_sustained'], 'traffic': 0.922e9 * last_mb * (S / 512) ** 2, 'traffic_note': 'dram__bytes_read+write summed over the 11 launches of a 32-image pass from the ncu --set full capture in profiles/r01_ncu_halo_kernels.md (0.922 GB per 512x512 image), scaled to this pass', 'peak_source': pk['source'] + ' sustained bf16',
                     'tensor_pipe_active_pct_ncu': {'cout_ge_128_cta_pair': [73.8, 89.6], 'cout_64': [62.4, 71.3], 'upconv': [26.7, 39.9],
                                                    'source': 'profiles/r01_ncu_chain_final.md (ncu --set full, sm__pipe_tensor_cycles_active, not measured in this run)'},
                     'kernel': 'conv_halo_kernel / conv_halo2_kernel / upconv_res_kernel (11 tensor-core launches per micro-batch)', 'issued_tflops': 3 * achieved,
                     'issued_frac': 3 * achieved / pk['tf_sustained'],
                     'note': 'achieved = algorithmic 2*MACs of the 11 tensor-core layers / their summed CUDA-event time in the last micro-batch of the timed region; '
                             'every MAC is issued as 3 bf16 MMAs (hi*hi, lo*hi, hi*lo), so issued = 3x algorithmic'},
        'layers': layers,
        'estimator': est,
        'cpu_baseline': cpu_base,
    }
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--per-gpu', type=int, default=256, help='images per GPU per step')
    ap.add_argument('--micro-batch', type=int, default=0)
    ap.add_argument('--est-images', type=int, default=10000, help='images of the KB-filter estimator measurement (configs[1]); 0 skips it (launch lists of the UNet step)')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--size', type=int, default=512, help='image side; 1024 = BASELINE configs[4] (not the headline metric)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
