/*
 * NVText — tokenize / token_count of the reference's cpp/include/NVText.h:29-174 on libcustr.so's C-ABI.
 */
#pragma once
#include "NVStrings.h"

class NVText {
public:
    static NVStrings* tokenize(NVStrings& strs, const char* delimiter = nullptr);                                          // :40
    static unsigned int token_count(NVStrings& strs, const char* delimiter, unsigned int* results, bool devmem = true);   // :66
};
