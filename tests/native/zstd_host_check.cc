// Host build of paimon_b200/csrc/zstd_device.cuh (the same source the device kernel compiles): a C entry point for
// tests/test_zstd_cpu.py, which pins the decoder against pyarrow / libzstd-compressed buffers without a GPU.
#include <stdlib.h>

#include "zstd_device.cuh"

extern "C" long long zs_host_decode(const unsigned char *src, long long n, unsigned char *dst, long long cap) {
    zs::Tables *T = (zs::Tables *)calloc(1, sizeof(zs::Tables));
    unsigned char *lit = (unsigned char *)malloc(zs::kMaxBlock + 64);
    const long long r = zs::decode(src, n, dst, cap, lit, *T);
    free(lit);
    free(T);
    return r;
}
