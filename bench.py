#!/usr/bin/env python
"""Benchmark of the UNet -> WS hot path (BASELINE.json metric: UNet-WS 512x512 images/sec; conv tensor-pipe
fraction; estimator HBM fraction).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--images TOTAL] [--size 512|1024]

One "step" = one pass of the hot path over one batch of synthetic input: PER_GPU (default 256) 512x512 uint8 LSBr
(alpha sweep) stego images per GPU through the fused UNet(unet_2, random init) -> WS beta_hat chain
(BASELINE.json configs[2]). At N > 1 rank r owns images [r*PER_GPU, (r+1)*PER_GPU) of the global index space
(generated on its GPU from per-chunk seeds, so any rank can reproduce any other rank's images) = configs[3], weak
scaling, no data-path collective, one NCCL all_gather of beta_hat per step; rank 0 recomputes a slice owned by another
rank and reports whether the gathered values equal it bit for bit. `--images TOTAL` runs the same thing with
TOTAL / N images per rank per step (configs[3] proper: --images 100000 --steps 1).
`value` is images/s with inputs resident in HBM; `e2e` is the same metric through the host-buffer C-ABI call (pinned
host images -> H2D -> chain -> D2H of beta_hat/l1 inside the timed region).
`--impl reference` times the reference's own CPU implementation of the same per-image path on the host cores: the
UNMODIFIED reference package mirrored into baseline/_ref/src by __graft_entry__.build() (kind "reference"), or the
ATen port oracle/torch_port.py when that copy is absent (kind "port").
"""
import argparse
import ctypes
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMG_512 = 202.199          # algorithmic 2*MACs of unet_2 at 512x512 (exact)
LAYER_GFLOP = {'e11': 0.302, 'e12': 19.327, 'e21': 9.664, 'e22': 19.327, 'e31': 9.664, 'e32': 19.327, 'upconv3': 4.295,
               'd31': 38.655, 'd32': 19.327, 'upconv4': 4.295, 'd41': 38.655, 'd42': 19.327 + 0.034}
ALPHAS = [0.01, 0.05, 0.1, 0.2, 0.4, 1.0]
METRIC = 'UNet-WS 512x512 images/sec'
MODEL_SEED = 1234
PX_TOL, BETA_TOL = 1e-3, 1e-4        # BASELINE.json north_star


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tf_burst': p['bf16_tflops'], 'tf_sustained': p['bf16_tflops_sustained'], 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'source': 'fallback'}


def ncu_summary():
    """Newest committed profiles/r*_ncu_chain.json (written by tools/ncu_chain_summary.py from an `ncu --set full` capture):
    DRAM traffic per image and tensor-pipe activity per layer. Evidence from a profiler run, labelled as such; null when no
    capture of the current kernels has been committed."""
    files = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'r*_ncu_chain.json')))
    if not files:
        return None
    try:
        d = json.load(open(files[-1]))
        d['file'] = os.path.relpath(files[-1], ROOT)
        return d
    except Exception:
        return None


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
        sm, mx, reasons, pw = [], None, set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm),
                'power_w_max': max(pw) if pw else None}


def make_inputs(first, n, device, size=512):
    """Images [first, first+n) of the global synthetic index space (ws_unet_b200/data.py::synthetic_stego_chunk)."""
    from ws_unet_b200 import data as wdata
    return wdata.synthetic_stego_shard(first, n, ALPHAS, size, size, device)


def build_model(device, seed=MODEL_SEED):
    import torch
    import ws_unet_b200 as W
    torch.manual_seed(seed)
    return W.get_model('unet_2', in_channels=1, out_channels=1, channel=[0], drop_rate=0.).to(device)


# ------------------------------------------------------------------------------------------------ CPU baseline
class CpuReference:
    """The reference's per-image UNet-WS path on the host cores (batch 1, FP32, autograd on, as src/unet/evaluate.py runs it)."""

    def __init__(self):
        import torch
        from oracle import reference_import
        self.torch = torch
        self.kind = 'port'
        if reference_import.reference_src() is not None:
            try:
                _defs, _f, runet, _ws = reference_import.import_reference()
                torch.manual_seed(MODEL_SEED)
                self.model = runet.model.get_model('unet_2', in_channels=1, out_channels=1, channel=[0], drop_rate=0.)
                self.runet = runet
                self.kind = 'reference'
                self.what = ('unmodified reference package (baseline/_ref/src): unet.evaluate.predict_unet -> infere_single -> '
                             'UNet.forward, src/unet/evaluate.py:31-52,109-139')
            except Exception as e:   # missing third-party module on this box: fall back to the ATen port
                sys.stderr.write(f'[bench] reference import failed ({e!r}); using oracle/torch_port.py\n')
        if self.kind == 'port':
            import ws_unet_b200 as W
            torch.manual_seed(MODEL_SEED)
            self.model = W.get_model('unet_2', 1, 1, [0], 0.)
            self.what = 'oracle/torch_port.py: the torch ATen CPU ops exactly as src/unet/evaluate.py:31-52,125-132 calls them'

    def predict(self, img_u8, want_xhat=False):
        """beta_hat, l1 (and the 510x510 prediction) of one uint8 (H,W) image."""
        import numpy as np
        if self.kind == 'reference':
            x4 = np.repeat(img_u8[..., None], 4, axis=2).astype('float32')
            r = self.runet.evaluate.predict_unet('mem', self.model, imread=lambda f: x4)
            xh = self.runet.infere_single(x4[..., 3:], self.model)[..., 0] if want_xhat else None
            return float(r['beta_hat']), float(r['l1']), xh
        from oracle import torch_port
        b, l = torch_port.predict_unet(img_u8, self.model)
        xh = torch_port.infere_single(img_u8.astype('float32')[..., None], self.model, no_grad=True)[..., 0] if want_xhat else None
        return float(b), float(l), xh


def cpu_images(n, size=512):
    from ws_unet_b200 import data as wdata
    return [wdata.embed_lsbr(wdata.synthetic_cover(i, size, size), ALPHAS[(i + 4) % 6], i).numpy() for i in range(n)]   # 0.4, 1.0, ...


def cpu_baseline_run(ref, imgs, seconds=15.0, min_images=3):
    ref.predict(imgs[0])  # warm-up (first call pays oneDNN primitive creation)
    t0 = time.perf_counter()
    n, out = 0, {}
    while n < min_images or time.perf_counter() - t0 < seconds:
        i = n % len(imgs)
        out[i] = ref.predict(imgs[i])[:2]
        n += 1
        if n >= 64:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, out


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    per_step = 2
    ref = CpuReference()
    imgs = cpu_images(per_step)
    for _ in range(args.warmup):
        ref.predict(imgs[0])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for im in imgs:
            ref.predict(im)
    dt = time.perf_counter() - t0
    v = args.steps * per_step / dt
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'UNet-WS (unet_2 random init) on 512x512 LSBr alpha-sweep stego images, reference CPU path: '
                               'batch 1, FP32, autograd on (a per-image rate: the reference has no batched path)',
                   'images_per_step': per_step},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': ref.kind,
                         'sample': f'{args.steps} steps x {per_step} images; {ref.what}'},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def measure_estimator(W, torch, imgs, n_est, S, pk, local):
    """KB-filter WS on resident uint8 images (HBM-bound, BASELINE.json configs[1]) + the same from pinned host memory."""
    est_sampler = ClockSampler(local)
    est_sampler.start()
    per = imgs.shape[0]
    est_imgs = imgs.repeat((n_est + per - 1) // per, 1, 1, 1)[:n_est].contiguous()
    est = {}
    # nvidia-smi takes ~1 s to come up and holds driver locks while it does: launches stall behind it, which a
    # 0.4 ms kernel shows. Warm up until its first sample has arrived, then time.
    t_wait = time.perf_counter()
    while not est_sampler.rows and time.perf_counter() - t_wait < 5.0:
        W.ws_estimate(est_imgs, 'KB', weighted=0)
        torch.cuda.synchronize()
    modes = [('kb_w0', dict(weighted=0), 'beta_hat, unweighted'), ('kb_w1', dict(weighted=1), 'beta_hat, 1/(5+var) weights (attack default)'),
             ('kb_w0_l1', dict(weighted=0, return_l1=True), 'beta_hat + L1, unweighted')]
    for key, kw, what in modes:
        for _ in range(10):   # also lets the clocks ramp before the timed repetitions
            W.ws_estimate(est_imgs, 'KB', **kw)
        torch.cuda.synchronize()
        groups, reps = [], 100   # long groups: the host runs ahead of the GPU and absorbs nvidia-smi polling stalls
        for _ in range(3):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                W.ws_estimate(est_imgs, 'KB', **kw)
            a1.record()
            torch.cuda.synchronize()
            groups.append(a0.elapsed_time(a1) / reps / 1e3)
        sec = sorted(groups)[1]   # median of three groups of 100 back-to-back calls
        gbs = (S * S + 4) * n_est / sec / 1e9
        est[key] = {'what': what, 'images_per_s': n_est / sec, 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                    'frac': gbs / pk['hbm_gbs'], 'bound': 'hbm', 'images': n_est}
    # the same estimator end to end from pinned host memory (H2D copy inside the timed region): host-link bound.
    # A plain pinned cudaMemcpy of the same bytes is timed next to it as the ceiling of this box's link.
    n_h = min(n_est, 4096)
    host_est = est_imgs[:n_h].cpu().pin_memory()
    dst = torch.empty_like(est_imgs[:n_h])
    # Both are short (about 20 ms per call) and run while nvidia-smi polls the clocks every 100 ms, which can stall a driver
    # call for tens of milliseconds: each is the best of 8 calls (the median is reported next to it).
    def best_of(fn, reps=8):
        ts = []
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return ts[0], ts[len(ts) // 2]

    copy_best, copy_med = best_of(lambda: dst.copy_(host_est, non_blocking=True))
    copy_gbs = S * S * n_h / copy_best / 1e9
    dt, dt_med = best_of(lambda: W.ws_estimate_host(host_est, 'KB', weighted=0))
    est['kb_w0_e2e'] = {'images_per_s': n_h / dt, 'h2d_gbs': S * S * n_h / dt / 1e9, 'pinned_copy_ceiling_gbs': copy_gbs, 'images': n_h,
                        'median_h2d_gbs': S * S * n_h / dt_med / 1e9, 'median_ceiling_gbs': S * S * n_h / copy_med / 1e9,
                        'note': 'wsu_filter_ws_estimate_host: pinned host uint8 -> H2D -> kernel -> D2H, best of 8 calls; ceiling = torch pinned->device copy of the same bytes on this box, best of 8'}
    del est_imgs, host_est, dst
    est['clocks'] = est_sampler.stop()
    torch.cuda.empty_cache()
    return est


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import ws_unet_b200 as W
    from ws_unet_b200 import _native, parallel
    from ws_unet_b200 import data as wdata

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
    G = wdata.GEN_CHUNK
    per_gpu = args.per_gpu
    if args.images:
        per_gpu = -(-args.images // (world * G)) * G     # whole generator chunks per rank
    if per_gpu % G:
        raise SystemExit(f'--per-gpu must be a multiple of {G} (generator chunk)')
    model = build_model(dev)
    S = args.size
    imgs = make_inputs(rank * per_gpu, per_gpu, dev, S)
    if args.micro_batch:
        model.set_micro_batch(args.micro_batch, dev)
    n_total = per_gpu * world
    h = model.native_handle(dev)
    precision = configure_precision(args, W, lib, h, model, imgs, dev)

    # measured first, as its own workload, before the tensor-core chain heats the part up
    est = None
    if rank == 0 and args.est_images > 0:
        est = measure_estimator(W, torch, imgs[:min(per_gpu, 256)], args.est_images, S, pk, local)

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
    per_rank_ms = [ms / args.steps]
    if world > 1:
        allms = torch.empty(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allms, t)
        per_rank_ms = [float(v) / args.steps for v in allms.tolist()]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    value = n_total * args.steps / (ms_max / 1e3)

    # ---- end to end through the host-buffer C-ABI call (pinned host memory -> H2D -> chain -> D2H)
    host_img = imgs.cpu().pin_memory()
    host_out = torch.empty(2, per_gpu, dtype=torch.float32).pin_memory()

    def e2e_step():
        _native.check(lib.wsu_unet_ws_estimate_host(h, ctypes.c_void_p(host_img.data_ptr()), per_gpu, S, S, 0, 1, 1, 0,
                                                    ctypes.c_void_p(host_out[0].data_ptr()), ctypes.c_void_p(host_out[1].data_ptr())))

    e2e_steps = args.steps if not args.images else 1
    for _ in range(max(1, args.warmup // 2) if not args.images else 0):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_total * e2e_steps / t.item()
    mine = out[rank * per_gpu:(rank + 1) * per_gpu] if world > 1 else out
    e2e_ok = bool(torch.equal(host_out[0].to(dev), mine))
    del host_img

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- N > 1: the gathered vector's slice owned by ANOTHER rank, recomputed here from the same chunk seeds
    gather_check = None
    if world > 1:
        other = 1
        n_chk = min(per_gpu, 64)
        theirs = make_inputs(other * per_gpu, n_chk, dev, S)
        again = W.ws_estimate(theirs, model, weighted=0, clip=True, crop=1)
        gather_check = {'bit_equal': bool(torch.equal(again, out[other * per_gpu: other * per_gpu + n_chk])),
                        'images': n_chk, 'owner_rank': other,
                        'differs_from_rank0_slice': bool(not torch.equal(out[:n_chk], out[other * per_gpu: other * per_gpu + n_chk]))}
        del theirs

    # ---- per-layer device times of the last micro-batch of the timed region -> roofline of the tensor-core chain
    info = ctypes.c_int64()
    lib.wsu_get_info(h, b'micro_batch', ctypes.byref(info))
    mb = int(info.value)
    scale = (S / 512) ** 2
    layers = []
    for i in range(max(0, n_layers)):
        name = lib.wsu_profile_name(h, i).decode()
        tf = LAYER_GFLOP.get(name, 0.0) * scale * last_mb / (prof_buf[i] * 1e-3) / 1e3 if prof_buf[i] > 0 else 0.0
        layers.append({'layer': name, 'ms': round(prof_buf[i], 4), 'tflops': round(tf, 1)})
    conv_ms = sum(l['ms'] for l in layers if l['layer'] != 'e11')
    conv_gflop = sum(LAYER_GFLOP[l['layer']] for l in layers if l['layer'] != 'e11') * scale
    achieved = conv_gflop * last_mb / (conv_ms * 1e-3) / 1e3 if conv_ms else 0.0
    terms = precision['terms_per_layer']
    issued = sum(LAYER_GFLOP[l['layer']] * terms.get(l['layer'], 3) for l in layers if l['layer'] != 'e11') * scale * last_mb / (conv_ms * 1e-3) / 1e3 if conv_ms else 0.0
    step_tflops = GFLOP_PER_IMG_512 * scale * value / world / 1e3      # per GPU, every launch of the step included
    top = max((l for l in layers if l['layer'] != 'e11'), key=lambda l: l['ms'], default=None)

    # ---- CPU baseline (reference's per-image path on this box's host cores), bounded sample; N = 1 only: under torchrun
    # every rank is pinned to one OpenMP thread, which would not be the reference's configuration. Its beta_hat / l1 and
    # one full prediction map are compared with the GPU path on the same images: parity checked in the same run.
    cpu_base, parity = None, None
    if world == 1 and args.cpu_seconds > 0 and S == 512:
        ref = CpuReference()
        cimgs = cpu_images(2)
        cpu_v, cpu_n, ref_vals = cpu_baseline_run(ref, cimgs, seconds=args.cpu_seconds)
        cores = torch.get_num_threads()
        cpu_base = {'value': cpu_v, 'unit': 'images/s', 'cores': cores, 'kind': ref.kind,
                    'sample': f'{cpu_n} images 512x512, batch 1, FP32, autograd on; {ref.what}'}
        dimg = torch.from_numpy(np.stack(cimgs))[:, None].to(dev)
        gb, gl, gy = W.ws_estimate(dimg, model, weighted=0, clip=False, crop=1, return_l1=True, return_prediction=True)
        xh_ref = ref.predict(cimgs[0], want_xhat=True)[2]
        d_px = float(np.abs((gy[0, 0, 1:-1, 1:-1] * 255.).cpu().numpy() - xh_ref).max())
        d_beta = max(abs(gb[i].item() - ref_vals[i][0]) for i in ref_vals)
        d_l1 = max(abs(gl[i].item() - ref_vals[i][1]) for i in ref_vals)
        parity = {'max_abs_beta': d_beta, 'max_abs_px': d_px, 'max_abs_l1': d_l1, 'images': len(ref_vals), 'against': ref.kind,
                  'ok': bool(d_beta < BETA_TOL and d_px < PX_TOL), 'tolerance': {'beta': BETA_TOL, 'px': PX_TOL}}

    # ---- the same step under the other precision plans (N = 1 only, after everything that is reported above): what the
    # headline would be with the round-1 arithmetic (three MMAs per MAC everywhere) and with the intermediate plan
    plans = None
    if world == 1 and S == 512 and not args.images and args.precision == 'auto':
        plans = {}
        chosen = model.active_precision(dev)
        for mode in ('bf16x3', 'fp16x1', 'fp16x1_f8'):
            if mode == chosen:
                plans[mode] = {'images_per_s': value, 'note': 'the timed region above'}
                continue
            model.set_precision(mode)
            if model.active_precision(dev) != mode:
                continue
            for _ in range(2):
                W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
            torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record()
            for _ in range(3):
                W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
            p1.record()
            torch.cuda.synchronize()
            plans[mode] = {'images_per_s': per_gpu * 3 / (p0.elapsed_time(p1) / 1e3), 'note': '3 steps after the timed region (secondary)'}
        model.set_precision(chosen)

    ncu = ncu_summary()
    line = {
        'metric': METRIC if S == 512 else f'UNet-WS {S}x{S} images/sec', 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': precision['dtype'], 'data': 'synthetic',
        'config': {'workload': f'UNet-WS (unet_2 random init) beta_hat on {per_gpu} synthetic {S}x{S} uint8 LSBr stego images per '
                               f'GPU, alpha sweep {ALPHAS}, weighted=0 (BASELINE configs[2]; at N>1 rank r owns global images '
                               f'[r*{per_gpu}, (r+1)*{per_gpu}) = configs[3], {n_total} images per step)',
                   'images_per_step': n_total, 'micro_batch': mb, 'parallelism': f'image-sharded x{world}, all_gather(beta_hat)',
                   'arithmetic': precision['arithmetic'],
                   'l2_policy': 'working set per micro-batch (>= 5 GB of activations) far exceeds the 126 MB L2; no flush needed'},
        'clocks': clocks,
        'per_rank_ms': [round(v, 3) for v in per_rank_ms],
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': per_gpu * S * S * world,
                'd2h_bytes_per_step': per_gpu * 8 * world, 'matches_device_path': e2e_ok},
        'gpu_launches': int(launches),
        'gather_check': gather_check,
        'parity': parity,
        'precision': precision['report'],
        'precision_plans': plans,
        'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                     'frac': achieved / pk['tf_sustained'],
                     'frac_step': step_tflops / pk['tf_sustained'],
                     'frac_step_note': f'{GFLOP_PER_IMG_512} GFLOP/image x value / n_gpus / peak: every launch of the step (e11, finalize, gather) included',
                     'traffic': (ncu['dram_bytes_per_image'] * last_mb * scale) if ncu and ncu.get('dram_bytes_per_image') else None,
                     'traffic_note': (f"dram__bytes_read+write per image from the ncu --set full capture summarised in {ncu['file']} "
                                      f"({ncu.get('captured', '')}), scaled to this pass of {last_mb} images; not measured in this run") if ncu else
                                     'no ncu capture of the current kernels committed',
                     'tensor_pipe_active_pct_ncu': ncu.get('tensor_active_pct') if ncu else None,
                     'peak_source': pk['source'] + ' sustained bf16',
                     'kernel': 'conv_halo_kernel / conv_halo2_kernel / upconv_res_kernel (11 tensor-core launches per micro-batch)',
                     'dominant_layer': top,
                     'issued_tflops': issued, 'issued_frac': issued / pk['tf_sustained'],
                     'note': 'achieved = algorithmic 2*MACs of the 11 tensor-core layers / their summed CUDA-event time in the last micro-batch of the timed region; '
                             'issued = the same with every MAC counted once per MMA term actually issued for its layer (precision.terms_per_layer)'},
        'layers': layers,
        'estimator': est,
        'cpu_baseline': cpu_base,
    }
    os.write(json_fd, (json.dumps(line) + '\n').encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def configure_precision(args, W, lib, h, model, imgs, dev):
    """Precision plan of the layers whose input lives at UNet level >= 1 (ws_unet_b200/unet/model.py::set_precision).
    'auto' (default): UNet.calibrate_precision on 8 images of this rank's shard - the cheapest plan whose predictions stay
    within 2e-4 px (max-abs over all pixels) of the three-term plan is used; the in-run `parity` block then checks the
    chosen plan against the CPU reference like any other."""
    names = ['e12', 'e21', 'e22', 'e31', 'e32', 'upconv3', 'd31', 'd32', 'upconv4', 'd41', 'd42']
    deep = ['e21', 'e22', 'e31', 'e32', 'upconv3', 'd31', 'd32', 'upconv4']
    if args.precision == 'auto':
        report = model.calibrate_precision(imgs[:8], budget_px=2e-4)
    else:
        model.set_precision(args.precision)
        report = {'chosen': model.active_precision(dev), 'requested': args.precision}
    mode = model.active_precision(dev)
    t = {'bf16x3': 3, 'fp16x2': 2, 'fp16x1': 1, 'fp16x1_f8': 1}[mode]
    terms = {n: (t if n in deep else 3) for n in names}
    if mode in ('fp16x2', 'fp16x1'):
        terms['d41'] = 2.5   # its upsampled half (u4, fp16) takes two MMAs per MAC, the skip half (e12) three
    if mode == 'fp16x1_f8':  # full-resolution layers: one fp16 MMA + one e4m3 MMA (K = 32 per instruction) per MAC
        terms.update({'e12': 2, 'd42': 2, 'd41': 1.5})
    if mode == 'bf16x3':
        arith = 'split-bf16 operands (hi + lo), 3 tcgen05 MMAs per MAC, fp32 accumulation in TMEM; integer WS arithmetic'
        dtype = 'bf16x3'
    elif mode == 'fp16x1_f8':
        arith = ('full-resolution layers e12/d41/d42: fp16 main product + ONE e4m3 MMA over both correction terms (2 MMA times per MAC, '
                 '~15 significant bits); layers at level >= 1: one fp16 activation plane x fp16 weights, 1 MMA per MAC (chosen by '
                 'calibration against the 3-term plan); fp32 accumulation in TMEM; integer WS arithmetic')
        dtype = 'fp16+e4m3x2 / fp16x1'
    else:
        arith = (f'full-resolution layers e12/d41/d42: split-bf16 operands, 3 tcgen05 MMAs per MAC; layers at level >= 1: one fp16 '
                 f'activation plane x fp16 weights, {t} MMA(s) per MAC ({mode}, chosen by calibration against the 3-term plan); '
                 'fp32 accumulation in TMEM; integer WS arithmetic')
        dtype = f'bf16x3+{mode}'
    report['mode'] = mode
    return {'dtype': dtype, 'terms_per_layer': terms, 'arithmetic': arith, 'report': report}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--per-gpu', type=int, default=256, help='images per GPU per step (multiple of 64)')
    ap.add_argument('--images', type=int, default=0, help='total images per step over all GPUs (configs[3]: 100000); overrides --per-gpu')
    ap.add_argument('--micro-batch', type=int, default=0)
    ap.add_argument('--est-images', type=int, default=10000, help='images of the KB-filter estimator measurement (configs[1]); 0 skips it (launch lists of the UNet step)')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--size', type=int, default=512, help='image side; 1024 = BASELINE configs[4] (not the headline metric)')
    ap.add_argument('--precision', default='auto', choices=['auto', 'bf16x3', 'fp16x2', 'fp16x1', 'fp16x1_f8'])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
