"""Dev tool: what the runtime says about co-residency of the conv kernel's two configurations."""
import ctypes
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from yolov5_obb_b200 import _lib

torch.zeros(1, device="cuda")
L = _lib.lib()
p = torch.cuda.get_device_properties(0)
print("regs/SM", p.regs_per_multiprocessor, "smem/SM", p.shared_memory_per_multiprocessor, "smem/block optin", p.shared_memory_per_block_optin)
for dual, th in ((1, 192), (1, 256), (0, 320)):
    for kb in (60, 80, 100, 104, 108, 110, 112, 116, 200):
        b, r, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = L.y5obb_conv_debug_occupancy(dual, th, kb * 1024, ctypes.byref(b), ctypes.byref(r), ctypes.byref(s))
        print(f"dual={dual} threads={th} dyn_smem={kb} KB -> rc={rc} blocks/SM={b.value} regs={r.value} static_smem={s.value}")
for kb in (60, 100, 110):
    b, r = ctypes.c_int(), ctypes.c_int()
    rc = L.y5obb_wgrad_debug_occupancy(kb * 1024, ctypes.byref(b), ctypes.byref(r))
    print(f"wgrad threads=192 dyn_smem={kb} KB -> rc={rc} blocks/SM={b.value} regs={r.value}")
