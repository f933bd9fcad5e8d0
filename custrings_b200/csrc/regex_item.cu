// Item-buffered chain kernel (regex_chain_item.cuh), compiled in four translation units (-DITEM_NS_GROUP=0..3: chains of
// 1-2, 3-4, 5-6, 7-8 steps) so that the instantiations build in parallel.  Host entry: launch_chain_item() in regex_bits.cu.
#include "regex_bits.h"
#include "regex_bits_plan.h"
#include "regex_vm.cuh"
#include <cstddef>
#include <mutex>
#include <unordered_map>

#ifndef ITEM_NS_GROUP
#define ITEM_NS_GROUP 0
#endif

namespace custr {
namespace bits {

#include "regex_bits_dev.cuh"
#define CUSTR_NO_LAUNCHERS
#include "regex_chain.cuh"
#include "regex_chain64.cuh"
#include "regex_chain_item.cuh"

#if ITEM_NS_GROUP == 1
int chain_item_ctas_per_sm() { return ITEM_MIN_CTAS; }  // resident CTAs per SM the kernels are built for (grid = SMs x this)
#endif
#define ITEM_ENTRY_NAME2(g) launch_chain_item_g##g
#define ITEM_ENTRY_NAME(g) ITEM_ENTRY_NAME2(g)
void ITEM_ENTRY_NAME(ITEM_NS_GROUP)(const ChainDev& cd, const Args& a, int blocks)
{
#ifdef ITEM_EXPERIMENT  // quick SASS iteration on the headline instantiation (tools/sass_stat.sh)
    auto kfn = k_chain_item<4, 1, ITEM_EXPERIMENT>;
    LAUNCH(kfn, blocks, THREADS, ITEM_SMEM_BYTES, cd, a);
    return;
#endif
    if ((int)cd.nsteps <= 2 * ITEM_NS_GROUP + 1) launch_item_ns<2 * ITEM_NS_GROUP + 1>(cd, a, blocks);
    else launch_item_ns<2 * ITEM_NS_GROUP + 2>(cd, a, blocks);
}

}  // namespace bits
}  // namespace custr

#if defined(CUSTR_ITEM_TIMING) && ITEM_NS_GROUP == 1
extern "C" int custr_dbg_item_times(unsigned long long* out, int max_words)
{
    const int n = 148 * 3 * custr::bits::WARPS * custr::bits::ITEM_TIMING_SLOTS;
    if (max_words < n) return -n;
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, custr::bits::g_item_times, sizeof(unsigned long long) * n);
    return n;
}
extern "C" void custr_dbg_item_times_clear(void)
{
    void* p = nullptr;
    cudaGetSymbolAddress(&p, custr::bits::g_item_times);
    cudaMemset(p, 0, sizeof(custr::bits::g_item_times));
}
#endif
