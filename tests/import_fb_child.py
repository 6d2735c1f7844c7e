"""Child process of tests/test_gpu_present.py: imports the framebuffer another process exported with
fdc_export_framebuffer (POSIX file descriptor inherited from the parent) through the CUDA driver API -- the same
steps a Vulkan / GL presenter performs with its own external-memory import -- and dumps the pixels.
usage: import_fb_child.py <fd> <bytes> <width> <height> <out.npy>"""
import sys

import numpy as np
from cuda.bindings import driver as cu


def ck(res):
    err, *rest = res
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(f"CUDA driver error {err}")
    return rest[0] if len(rest) == 1 else rest


fd, nbytes, w, h, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
ck(cu.cuInit(0))
dev = ck(cu.cuDeviceGet(0))
ctx = ck(cu.cuDevicePrimaryCtxRetain(dev))
ck(cu.cuCtxSetCurrent(ctx))
handle = ck(cu.cuMemImportFromShareableHandle(fd, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR))
va = ck(cu.cuMemAddressReserve(nbytes, 0, 0, 0))
ck(cu.cuMemMap(va, nbytes, 0, handle, 0))
acc = cu.CUmemAccessDesc()
acc.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
acc.location.id = 0
acc.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READ
ck(cu.cuMemSetAccess(va, nbytes, [acc], 1))
pix = np.empty((h, w, 4), dtype=np.uint8)
ck(cu.cuMemcpyDtoH(pix.ctypes.data, va, w * h * 4))
np.save(out, pix)
ck(cu.cuMemUnmap(va, nbytes))
ck(cu.cuMemRelease(handle))
ck(cu.cuMemAddressFree(va, nbytes))
print("imported", w, h)
