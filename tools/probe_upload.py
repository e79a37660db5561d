"""Upload path check: is the two-stream path taken for pinned uint16 views, and what does a pair cost?"""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from image_compression_analysis_b200 import engine
shape = (1024, 1024, 180)
r = torch.empty(shape, dtype=torch.int16).pin_memory(); d = torch.empty(shape, dtype=torch.int16).pin_memory()
r.zero_(); d.zero_()
ru, du = r.view(torch.uint16), d.view(torch.uint16)
print("pinned views:", ru.is_pinned(), du.is_pinned(), ru.dtype)
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("upload_pair (pinned, two streams) ms", t(lambda: engine.upload_pair(ru, du)))
print("to_device x2 (one stream)         ms", t(lambda: (engine.to_device(ru), engine.to_device(du))))
print("from_arrays                        ms", t(lambda: engine.DevicePair.from_arrays(ru, du, "bip")))
