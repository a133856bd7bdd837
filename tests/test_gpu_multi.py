"""GPU, needs >= 2 devices (skipped otherwise): BASELINE.json configs[3] in miniature - images sharded by contiguous index
range over one process per GPU, no data-path collective, one NCCL all_gather of beta_hat; the gathered vector must equal
the vector a single GPU computes, bit for bit."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
N_IMAGES, H, W = 21, 64, 96      # 21 images over 2 ranks: ragged shards (11 + 10)


def _images():
    from ws_unet_b200 import data as wdata
    return torch.stack([wdata.embed_lsbr(wdata.synthetic_cover(i, H, W), 0.4, i) for i in range(N_IMAGES)])[:, None]


def _nccl_worker(rank, world, port, q):
    import numpy as np
    import torch.distributed as dist
    import ws_unet_b200 as Wp
    from oracle import unet_oracle as uo
    from ws_unet_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_from_env('nccl')
    dev = torch.device('cuda', rank)
    imgs = _images()
    m = Wp.get_model('unet_2', 1).to(dev)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in uo.numpy_weights(2, seed=9).items()})
    load = lambda lo, hi: imgs[lo:hi].to(dev)
    out = {}
    out['unet'] = parallel.estimate_sharded(N_IMAGES, load, lambda x: Wp.ws_estimate(x, m, weighted=0, clip=False), chunk=4)
    out['kb'] = parallel.estimate_sharded(N_IMAGES, load, lambda x: Wp.ws_estimate(x, 'KB', weighted=0, clip=False), chunk=7)
    out['kb_w1'] = parallel.estimate_sharded(N_IMAGES, load, lambda x: Wp.ws_estimate(x, 'KB', weighted=1), chunk=5)
    torch.cuda.synchronize()
    q.put((rank, {k: v.cpu().numpy() for k, v in out.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_nccl_gather_equals_single_gpu():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices')
    import numpy as np
    import torch.multiprocessing as mp
    import ws_unet_b200 as Wp
    from oracle import unet_oracle as uo
    dev = torch.device('cuda', 0)
    imgs = _images().to(dev)
    m = Wp.get_model('unet_2', 1).to(dev)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in uo.numpy_weights(2, seed=9).items()})
    single = {'unet': Wp.ws_estimate(imgs, m, weighted=0, clip=False).cpu().numpy(),
              'kb': Wp.ws_estimate(imgs, 'KB', weighted=0, clip=False).cpu().numpy(),
              'kb_w1': Wp.ws_estimate(imgs, 'KB', weighted=1).cpu().numpy()}
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for _, got in res:                       # every rank holds the full gathered vector
        for key in single:
            assert np.array_equal(got[key], single[key]), key


def test_one_thread_two_devices_host_paths_and_current_device():
    """One thread driving two GPUs: the host-buffer entry points keep their staging streams per device, every C entry point
    leaves the caller's current device as it found it, and both devices give the same bits."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs two CUDA devices')
    import ws_unet_b200 as Wp
    from oracle import unet_oracle as uo
    imgs = _images()
    sd = {k: torch.from_numpy(v) for k, v in uo.numpy_weights(2, seed=9).items()}
    torch.cuda.set_device(0)
    res = {}
    for d in (0, 1, 0, 1):
        dev = torch.device('cuda', d)
        m = Wp.get_model('unet_2', 1).to(dev)
        m.load_state_dict(sd)
        kb = Wp.ws_estimate_host(imgs, 'KB', weighted=1, device=dev)
        un = Wp.ws_estimate_host(imgs, m, weighted=0, clip=False, device=dev)
        dv = Wp.ws_estimate(imgs.to(dev), m, weighted=0, clip=False).cpu()
        assert torch.cuda.current_device() == 0            # nobody switched the thread's device behind our back
        assert torch.equal(un, dv)
        res.setdefault('kb', kb)
        res.setdefault('un', un)
        assert torch.equal(kb, res['kb']) and torch.equal(un, res['un'])
        del m                                               # wsu_destroy on the other device must not switch it either
        assert torch.cuda.current_device() == 0
