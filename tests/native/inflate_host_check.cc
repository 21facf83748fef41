// Host build of paimon_b200/csrc/inflate_device.cuh (the same source the device kernels compile): C entry points for
// tests/test_inflate_cpu.py, which pins the decoder against zlib without a GPU.
#include <stdlib.h>

#include "inflate_device.cuh"

extern "C" long long if_host_raw(const unsigned char *src, long long n, unsigned char *dst, long long cap) {
    inflate::Tables *T = (inflate::Tables *)calloc(1, sizeof(inflate::Tables));
    const long long r = inflate::inflate_raw(src, n, dst, cap, *T, nullptr);
    free(T);
    return r;
}
extern "C" long long if_host_gzip(const unsigned char *src, long long n, unsigned char *dst, long long cap) {
    inflate::Tables *T = (inflate::Tables *)calloc(1, sizeof(inflate::Tables));
    const long long r = inflate::inflate_gzip(src, n, dst, cap, *T);
    free(T);
    return r;
}
extern "C" long long if_host_zlib(const unsigned char *src, long long n, unsigned char *dst, long long cap) {
    inflate::Tables *T = (inflate::Tables *)calloc(1, sizeof(inflate::Tables));
    const long long r = inflate::inflate_zlib(src, n, dst, cap, *T);
    free(T);
    return r;
}
