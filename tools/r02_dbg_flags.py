"""debug: stream flags on one device"""
import ctypes, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from innfer_b200 import _native as N
lib = N.load()
dev = torch.device("cuda:0")
flags = torch.zeros(8, dtype=torch.int32, device=dev)
marker = torch.zeros(1, dtype=torch.int32, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
torch.cuda.synchronize()
arr = (ctypes.c_void_p * 2)(flags.data_ptr(), flags.data_ptr() + 4)
t0 = time.time()
rc = lib.innfer_stream_wait(arr, 2, 5, ctypes.c_void_p(flags.data_ptr() + 28), 3000, ctypes.c_void_p(s1.cuda_stream))
print("wait rc", rc, "query right after", s1.query())
with torch.cuda.stream(s1):
    marker.add_(1)
print("query after add", s1.query())
time.sleep(0.2)
print("query after 0.2s", s1.query())
rc = lib.innfer_stream_signal(arr, 2, 5, ctypes.c_void_p(s2.cuda_stream))
s1.synchronize()
print("signal rc", rc, "released after %.3f s" % (time.time() - t0), "marker", marker.item(), "flags", flags.tolist())
arr1 = (ctypes.c_void_p * 1)(flags.data_ptr() + 8)
t0 = time.time()
lib.innfer_stream_wait(arr1, 1, 1, ctypes.c_void_p(flags.data_ptr() + 28), 50, ctypes.c_void_p(s1.cuda_stream))
s1.synchronize()
print("timeout wait took %.3f s" % (time.time() - t0), "flags", flags.tolist())
