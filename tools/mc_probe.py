"""Probe: is NVSwitch multicast (cuMulticast*) available on this box?"""
from cuda.bindings import driver as drv
def ck(r):
    if r[0] != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(r[0]))
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)
ck(drv.cuInit(0))
n = ck(drv.cuDeviceGetCount())
print("devices", n)
for d in range(n):
    dev = ck(drv.cuDeviceGet(d))
    mc = ck(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev))
    fab = ck(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, dev))
    fd = ck(drv.cuDeviceGetAttribute(drv.CUdevice_attribute.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev))
    print("dev", d, "multicast", mc, "fabric handles", fab, "posix fd", fd)
dev = ck(drv.cuDeviceGet(0))
ctx = ck(drv.cuDevicePrimaryCtxRetain(dev)); ck(drv.cuCtxSetCurrent(ctx))
prop = drv.CUmulticastObjectProp()
prop.numDevices = n
prop.size = 1 << 28
prop.handleTypes = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
prop.flags = 0
try:
    gran = ck(drv.cuMulticastGetGranularity(prop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED))
    print("granularity", gran)
    mcobj = ck(drv.cuMulticastCreate(prop))
    print("cuMulticastCreate ok", mcobj)
except Exception as e:
    print("multicast create failed:", e)
