// Small device helpers shared by the kernels: UTF-8 stepping, packed-char decode, unicode class flags.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define CUSTR_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define CUSTR_LDG(p) __ldg(p)
#else
#define CUSTR_LDG(p) (*(p))
#endif

namespace custr {

// width of the UTF-8 sequence introduced by lead byte b (reference custring_view.inl:48-57; a stray
// continuation byte yields 0 there, we step 1 so that scans always terminate)
CUSTR_HD int utf8_width(uint8_t b)
{
    int w = 1 + ((b & 0xF0) == 0xF0) + ((b & 0xE0) == 0xE0) + ((b & 0xC0) == 0xC0);
    return w;
}

// decode the packed char (bytes big-endian in a u32) at p, never reading at or past `end`
CUSTR_HD uint32_t utf8_packed(const uint8_t* p, const uint8_t* end, int& width)
{
    uint32_t c = *p;
    int w = 1;
    if (c >= 0xC0) {
        w = utf8_width((uint8_t)c);
        for (int k = 1; k < w; ++k) c = (c << 8) | (p + k < end ? p[k] : 0);
    }
    width = w;
    return c;
}

// packed char that ends right before `p` (p > begin)
CUSTR_HD uint32_t utf8_packed_before(const uint8_t* p, const uint8_t* begin)
{
    const uint8_t* q = p - 1;
    while (q > begin && (*q & 0xC0) == 0x80) --q;
    uint32_t c = 0;
    for (; q < p; ++q) c = (c << 8) | *q;
    return c;
}

CUSTR_HD uint32_t packed_to_cp(uint32_t c)  // reference util.inl:51-75
{
    if (c < 0x80u) return c;
    if (c < 0xE000u) return ((c & 0x1F00u) >> 2) | (c & 0x3Fu);
    if (c < 0xF00000u) return ((c & 0x0F0000u) >> 4) | ((c & 0x3F00u) >> 2) | (c & 0x3Fu);
    if (c <= 0xF8000000u) return ((c & 0x03000000u) >> 6) | ((c & 0x3F0000u) >> 4) | ((c & 0x3F00u) >> 2) | (c & 0x3Fu);
    return 0;
}

// alphanumeric in the sense of \b (reference regexec.inl:322-330): flags bits 0-3, code points <= 0xFFFF
CUSTR_HD bool is_alnum_packed(uint32_t c, const uint8_t* __restrict__ uflags)
{
    if (c < 0x80u) {
        uint32_t l = c | 0x20u;
        return (c - '0' < 10u) || (l - 'a' < 26u);
    }
    uint32_t cp = packed_to_cp(c);
    return cp < 0x10000u && (CUSTR_LDG(uflags + cp) & 15) != 0;
}

// number of characters in [p, p+n)
CUSTR_HD int utf8_count_chars(const uint8_t* p, int n)
{
    int cnt = 0;
    for (int i = 0; i < n; ++i) cnt += (p[i] & 0xC0) != 0x80;
    return cnt;
}

// byte offset of character index `pos` (clamped to n)
CUSTR_HD int utf8_offset_of(const uint8_t* p, int n, int pos)
{
    int off = 0;
    while (pos > 0 && off < n) {
        ++off;
        while (off < n && (p[off] & 0xC0) == 0x80) ++off;
        --pos;
    }
    return off;
}

const uint8_t* device_unicode_flags();  // 65536-entry table in HBM (uploaded once)
const uint8_t* host_unicode_flags();

}  // namespace custr
