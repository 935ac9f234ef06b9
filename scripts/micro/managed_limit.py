"""Diagnostic for multi-rank runs: how many 1 GiB managed blocks can each rank allocate, before and
after taking page-locked host memory?  torchrun --nproc-per-node N scripts/micro/managed_limit.py"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import vulkpy_b200 as vk
from vulkpy_b200 import _backend as B

rank = int(os.environ.get("LOCAL_RANK", 0))
gpu = vk.GPU(rank)
ctx = gpu.gpu._ctx
GiB = 1 << 30

def meminfo():
    d = {}
    for l in open("/proc/meminfo"):
        k, v = l.split(":")
        if k in ("MemFree", "CommitLimit", "Committed_AS", "Mlocked", "Unevictable"):
            d[k] = int(v.split()[0]) >> 20
    return d

def managed(n):
    got = []
    for i in range(n):
        p = C.c_void_p()
        rc = B.lib.vkp_alloc(ctx, GiB, C.byref(p))
        if rc != 0:
            return got, B.lib.vkp_last_error().decode()
        got.append(p.value)
    return got, None

log = [f"rank {rank} overcommit={open('/proc/sys/vm/overcommit_memory').read().strip()} {meminfo()}"]
a, err = managed(12)
log.append(f"  before pinned: {len(a)} managed GiB ok, err={err}")
pins = []
for i in range(4):
    p = C.c_void_p()
    rc = B.lib.vkp_host_alloc(GiB, C.byref(p))
    if rc != 0:
        log.append(f"  pinned {i} failed: {B.lib.vkp_last_error().decode()}")
        break
    pins.append(p.value)
    C.memset(p.value, 1, GiB)
b, err = managed(12)
log.append(f"  after {len(pins)} GiB pinned: {len(b)} more managed GiB ok, err={err} {meminfo()}")
gpu.wait()
time.sleep(2)
c, err = managed(4)
log.append(f"  2 s later: {len(c)} more, err={err}")
print("\n".join(log), flush=True)
