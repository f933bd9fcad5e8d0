/*
 * NVCategory — key-build subset of the reference's cpp/include/NVCategory.h:48-351 on libcustr.so's C-ABI.
 * keys = distinct strings sorted by unsigned-byte order (null first), values = int32 key index per row.
 */
#pragma once
#include <vector>
#include "NVStrings.h"

struct custr_category;

class NVCategory {
    custr_category* cat_;
    explicit NVCategory(custr_category* c) : cat_(c) {}
    ~NVCategory();
    NVCategory(const NVCategory&) = delete;

public:
    static NVCategory* create_from_array(const char** strs, unsigned int count);                        // :71
    static NVCategory* create_from_offsets(const char* strs, unsigned int count, const int* offsets,
                                           const unsigned char* nullbitmask = 0, int nulls = 0, bool devmem = true);  // :101
    static NVCategory* create_from_strings(NVStrings& strs);                                             // :107
    static NVCategory* create_from_strings(std::vector<NVStrings*>& strs);                               // :114
    static void destroy(NVCategory* inst);                                                               // :138

    unsigned int size();          // :148
    unsigned int keys_size();     // :152
    bool has_nulls();             // :156
    NVStrings* get_keys();        // :197
    int get_values(int* results, bool devmem = true);  // :225
    const int* values_cptr();     // :232
    NVStrings* to_strings();      // :326
    // category algebra (NVCategory.cu:430-514,1223-1345)
    static NVCategory* create_from_categories(std::vector<NVCategory*>& cats);   // :122  sorted union of keys, values remapped + appended
    NVCategory* merge_category(NVCategory& cat);                                 // :261  keys appended (not re-sorted)
    NVCategory* merge_and_remap(NVCategory& cat);                                // :270  = create_from_categories({this, cat})
    // key-set algebra and gathers (NVCategory.cu:1011-1220,1375-1820); positions are int32 key indexes
    NVCategory* add_keys_and_remap(NVStrings& strs);                             // :204
    NVCategory* remove_keys_and_remap(NVStrings& strs);                          // :213
    NVCategory* set_keys_and_remap(NVStrings& strs);                             // :234
    NVCategory* remove_unused_keys_and_remap();                                  // :242
    NVStrings* gather_strings(const int* pos, unsigned int elems, bool devmem = true);   // :291
    NVCategory* gather_and_remap(const int* pos, unsigned int elems, bool devmem = true); // :312
    NVCategory* gather(const int* pos, unsigned int elems, bool devmem = true);           // :335
};
