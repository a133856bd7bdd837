"""CPU: host-side mirror of the reference interfaces (module structure, state_dict, sharding, transforms)."""
import os
import pathlib
import pickle
import sys

import numpy as np
import pytest
import torch

import ws_unet_b200 as W
from ws_unet_b200 import parallel
from ws_unet_b200.unet.evaluate import _center_crop_512

REF = pathlib.Path('/root/reference/src/unet/model/unet.py')

PARAMS = {0: 37633, 1: 402625, 2: 1861697, 3: 7696193, 4: 31030593}


@pytest.mark.parametrize('nsteps', [0, 1, 2, 3, 4])
def test_parameter_count_and_keys(nsteps):
    m = W.get_model(f'unet_{nsteps}', in_channels=1, out_channels=1, channel=[0], drop_rate=0.)
    assert sum(p.numel() for p in m.parameters()) == PARAMS[nsteps]
    keys = set(m.state_dict())
    assert {'e11.weight', 'e11.bias', 'e12.weight', 'outconv.weight', 'outconv.bias'} <= keys
    if nsteps >= 2:
        assert m.state_dict()['upconv3.weight'].shape == (256, 128, 2, 2)
        assert m.state_dict()['d31.weight'].shape == (128, 256, 3, 3)
    assert len(list(m.buffers())) == 0


def test_get_model_errors_like_reference():
    with pytest.raises(NotImplementedError):
        W.get_model('resnet', 1)
    m = W.get_model('unet_2', 1)
    assert m.to('cpu') is m                       # unet.py:191-194 returns self
    m.disable_center_pixels()
    assert torch.all(m.e11.weight[:, :, 1, 1] == 0)
    m.input_dropout = None                        # saliency.py:140 assigns None
    m2 = pickle.loads(pickle.dumps(m))            # joblib workers pickle the closure (ws/estimate.py:139-146)
    assert m2._handle is None and torch.equal(m2.e12.weight, m.e12.weight)


@pytest.mark.skipif(not REF.exists(), reason='reference mount absent')
@pytest.mark.parametrize('nsteps', [0, 2])
def test_same_init_and_state_dict_as_reference(nsteps):
    import importlib.util
    spec = importlib.util.spec_from_file_location('ref_unet', REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(1234)
    a = ref.UNet(in_channels=1, out_channels=1, nsteps=nsteps, drop_rate=0., drop_channel=[0])
    torch.manual_seed(1234)
    b = W.get_model(f'unet_{nsteps}', 1, 1, [0], 0.)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    b.load_state_dict(sa)                          # interchangeable checkpoints (unet/evaluate.py:184-186)
    a.load_state_dict(sb)


def test_center_crop_matches_torchvision():
    import torchvision.transforms as T
    for h, w in [(512, 512), (600, 530), (500, 512), (513, 511)]:
        x = torch.rand(1, h, w)
        assert torch.equal(_center_crop_512(x), T.CenterCrop(512)(x)), (h, w)


def test_shard_range_partitions():
    for n in [0, 1, 7, 64, 100000]:
        for world in [1, 2, 3, 4, 8]:
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env('gloo')
    full = torch.arange(n, dtype=torch.float32) * 0.5 + 1

    def load(lo, hi):
        return full[lo:hi]

    out = parallel.estimate_sharded(n, load, lambda x: x * 2, chunk=3)
    q.put((r, out.tolist()))
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize('n', [10, 7])
def test_sharded_gather_world2_gloo(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = ((torch.arange(n, dtype=torch.float32) * 0.5 + 1) * 2).tolist()
    for _, got in res:
        assert got == expect


def test_processors_and_image_readers_match_reference_golden(defs_golden):
    """get_processor (N x 9 neighbour matrix, every inbayer lattice), get_processor_2d and the
    imread helpers reproduce the reference's outputs (tests/golden/make_golden_defs.py) exactly."""
    from ws_unet_b200 import defs
    img = defs_golden['proc_img']
    for inb in (None, '00', '01', '10', '11'):
        got = defs.get_processor(channels=(3,), inbayer=inb)(img)
        ref = defs_golden[f'proc_{inb}']
        assert got.shape == ref.shape and got.dtype == ref.dtype and np.array_equal(got, ref), inb
    for key, ch in (('proc2d_3', (3,)), ('proc2d_02', (0, 2))):
        got = defs.get_processor_2d(channels=ch)(img)
        assert got.dtype == np.float32 and np.array_equal(got, defs_golden[key])
    png = pathlib.Path(__file__).parent / 'golden' / 'rgb_9x11.png'
    x4 = defs.imread4_u8(png)
    assert x4.dtype == np.uint8 and np.array_equal(x4, defs_golden['imread4_rgb'])
    assert np.array_equal(defs.imread_u8(png), defs_golden['imread_u8_rgb'])
    assert defs.imread4_f32(png).dtype == np.float32 and defs.imread_f32(png).dtype == np.float32
    # the target column of the matrix is the interior of the image, the first column its top-left neighbours
    mat = defs.get_processor(channels=(3,))(img)
    assert np.array_equal(mat[:, 8], img[1:-1, 1:-1, 3].reshape(-1)) and np.array_equal(mat[:, 0], img[:-2, :-2, 3].reshape(-1))
