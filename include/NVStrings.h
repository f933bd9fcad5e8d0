/*
 * NVStrings — C++ class surface of the hot path, drop-in for the reference's cpp/include/NVStrings.h:48-1205
 * (same method names, argument meaning, return values and exception types), implemented on libcustr.so's C-ABI
 * (include/custr.h).  Instances are immutable; factories return new-ed objects released with NVStrings::destroy;
 * result arrays are caller allocated (device memory when devmem == true).  Only the methods of SURVEY.md §8 are
 * declared; everything else of the reference class is out of scope.
 */
#pragma once
#include <cstddef>
#include <utility>
#include <vector>

struct custr_column;

class NVStrings {
    custr_column* col_;
    explicit NVStrings(custr_column* c) : col_(c) {}
    ~NVStrings();
    NVStrings(const NVStrings&) = delete;
    NVStrings& operator=(const NVStrings&) = delete;
    friend class NVCategory;
    friend class NVText;

public:
    enum sorttype { none = 0, length = 1, name = 2 };  // reference NVStrings.h:55-59 (may be OR-ed)
    // reference NVStrings.h:86,98,116 / NVStrings.cu:74-119
    static NVStrings* create_from_array(const char** strs, unsigned int count);
    // (pointer, byte length) pairs; the pointers address device memory, devmem says where the pair array lives (:98)
    static NVStrings* create_from_index(std::pair<const char*, size_t>* strs, unsigned int count, bool devmem = true, sorttype stype = none);
    static NVStrings* create_from_offsets(const char* strs, int count, const int* offsets, const unsigned char* nullbitmask = 0,
                                          int nulls = 0, bool devmem = true);
    static void destroy(NVStrings* inst);  // :156

    unsigned int size() const;                                                                           // :167
    int create_offsets(char* strs, int* offsets, unsigned char* nullbitmask = 0, bool devmem = true);   // :207
    unsigned int set_null_bitarray(unsigned char* bitarray, bool emptyIsNull = false, bool devmem = true);  // :225
    int to_host(char** list, int start, int end);                                                       // :251
    unsigned int len(int* lengths, bool devmem = true);                                                 // :345
    size_t byte_count(int* lengths, bool devmem = true);                                                // :354
    int hash(unsigned int* results, bool devmem = true);                                                // :1031

    // regex: count.cu:59-250, replace.cu:110-189, replace_multi.cu:110-197
    int contains_re(const char* pattern, bool* results, bool devmem = true);                            // :963
    int match(const char* pattern, bool* results, bool devmem = true);                                  // :972
    int count_re(const char* pattern, int* results, bool devmem = true);                                // :981
    NVStrings* replace_re(const char* pattern, const char* repl, int maxrepl = -1);                     // :766
    NVStrings* replace_re(std::vector<const char*>& patterns, NVStrings& repls);                        // :779
    NVStrings* replace_with_backrefs(const char* pattern, const char* repl);     // :793  replace_backref.cu:122

    // capture-span callers ("next" rows): findall.cu:99, findall_record.cu:97, extract.cu:69
    int findall(const char* pattern, std::vector<NVStrings*>& results);          // :943  column c = c-th match of each row
    int findall_record(const char* pattern, std::vector<NVStrings*>& results);   // :952  one (possibly empty) instance per row
    int extract(const char* pattern, std::vector<NVStrings*>& results);          // :682  one column per capture group

    // literal: find.cu:75-387, modify.cu:109-299
    unsigned int find(const char* str, int start, int end, int* results, bool devmem = true);           // :861
    unsigned int rfind(const char* str, int start, int end, int* results, bool devmem = true);          // :873
    unsigned int find_from(const char* str, int* starts, int* ends, int* results, bool devmem = true);  // :885  (starts/ends live where devmem says)
    unsigned int find_multiple(NVStrings& strs, int* results, bool devmem = true);                      // :898
    int match_strings(NVStrings& strs, bool* results, bool devmem = true);                              // :916
    int contains(const char* str, bool* results, bool devmem = true);                                   // :907
    unsigned int startswith(const char* str, bool* results, bool devmem = true);                        // :925
    unsigned int endswith(const char* str, bool* results, bool devmem = true);                          // :934
    NVStrings* replace(const char* str, const char* repl, int maxrepl = -1);                            // :714
    NVStrings* replace(NVStrings& strs, NVStrings& repls);                                              // :726

    // split: split.cu:125-956.  split_record returns one NVStrings per row (nullptr for null rows); all of them are
    // views over ONE flat token column instead of the reference's N allocations.
    int split_record(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results);           // :464
    int split_record(int maxsplit, std::vector<NVStrings*>& results);                                   // :484
    unsigned int split(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results);          // :504
    unsigned int split(int maxsplit, std::vector<NVStrings*>& results);                                 // :524
    int rsplit_record(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results);          // :474
    int rsplit_record(int maxsplit, std::vector<NVStrings*>& results);                                  // :494
    unsigned int rsplit(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results);         // :514
    unsigned int rsplit(int maxsplit, std::vector<NVStrings*>& results);                                // :534
    // one instance of 3 strings per row (three nulls for a null row): split.cu:1165-1372
    int partition(const char* delimiter, std::vector<NVStrings*>& results);                             // :547
    int rpartition(const char* delimiter, std::vector<NVStrings*>& results);                            // :560

    NVStrings* gather(const int* pos, unsigned int count, bool devmem = true);                          // :270
    NVStrings* sublist(unsigned int start, unsigned int end, int step = 0);                             // :261

    custr_column* column() const { return col_; }  // escape hatch to the C-ABI handle

private:
    int record_split_(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results, bool right);
    unsigned int column_split_(const char* delimiter, int maxsplit, std::vector<NVStrings*>& results, bool right);
    int partition_(const char* delimiter, std::vector<NVStrings*>& results, bool right);

    // cheap attributes / transforms: attrs.cu:115-445, case.cu:30-190, strip.cu:30-200, substr.cu:32-83
    unsigned int isalnum(bool* results, bool todevice = true);
    unsigned int isalpha(bool* results, bool todevice = true);
    unsigned int isdigit(bool* results, bool todevice = true);
    unsigned int isspace(bool* results, bool todevice = true);
    unsigned int isdecimal(bool* results, bool todevice = true);
    unsigned int isnumeric(bool* results, bool todevice = true);
    unsigned int islower(bool* results, bool todevice = true);
    unsigned int isupper(bool* results, bool todevice = true);
    unsigned int is_empty(bool* results, bool todevice = true);
    NVStrings* lower();
    NVStrings* upper();
    NVStrings* strip(const char* to_strip = nullptr);
    NVStrings* lstrip(const char* to_strip = nullptr);
    NVStrings* rstrip(const char* to_strip = nullptr);
    NVStrings* slice(int start = 0, int stop = -1, int step = 1);
    NVStrings* get(unsigned int pos);
};
