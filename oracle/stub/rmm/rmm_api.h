#pragma once
#include <cstddef>
static inline int rmmGetInfo(size_t* f, size_t* t, void*) { *f = *t = (size_t)1 << 40; return 0; }
