"""One UNet-WS pass over N random 512x512 images (for ncu captures). Usage: python tools/one_pass.py [images=32] [precision=fp16x1] [passes=2]"""
import sys

import torch

sys.path.insert(0, '.')
import ws_unet_b200 as W

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else 'fp16x1'
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device('cuda', 0)
torch.manual_seed(1234)
model = W.get_model('unet_2', 1).to(dev).set_precision(mode)
imgs = torch.randint(0, 256, (n, 1, 512, 512), dtype=torch.uint8, device=dev)
model.set_micro_batch(n, dev)
for _ in range(passes):
    b = W.ws_estimate(imgs, model, weighted=0, clip=True, crop=1)
torch.cuda.synchronize()
print('ok', b[:2].tolist())
