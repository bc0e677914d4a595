import sys, io, time, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench, cv2
import os
import advmix_b200 as A
from advmix_b200 import _lib as _L
if os.environ.get('ADVMIX_B200_LIB'): _L.LIB_PATH = os.environ['ADVMIX_B200_LIB']
from advmix_b200 import jpeg as J
dev = torch.device('cuda:0'); B = 256
imgs = bench.natural_images_torch(B, dev, bench.SEED).cpu().numpy()
files = [cv2.imencode('.jpg', im, [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for im in imgs]
print('mean file bytes', np.mean([len(f) for f in files]))
enc = J.EncodedBatch(files)
for _ in range(2): sb = J.decode_batch(enc)
torch.cuda.synchronize()
t0 = time.perf_counter(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); a.record()
for _ in range(5): sb = J.decode_batch(enc)
b.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
print('decode_batch: %.3f ms gpu, %.3f ms wall per batch of %d -> %.0f img/s' % (a.elapsed_time(b) / 5, (t1 - t0) / 5 * 1e3, B, B / (a.elapsed_time(b) / 5e3)))
t0 = time.perf_counter()
for f in files[:64]: cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
print('cv2.imdecode 1 core: %.3f ms/img' % ((time.perf_counter() - t0) / 64 * 1e3))
