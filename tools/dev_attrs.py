import ctypes
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_int(0)
for name, attr in (("l2CacheSize", 38), ("maxPersistingL2CacheSize", 108), ("maxAccessPolicyWindowSize", 109), ("multiProcessorCount", 16),
                   ("maxSharedMemoryPerMultiprocessor", 81), ("maxRegistersPerMultiprocessor", 82)):
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0)
    print(name, v.value)
