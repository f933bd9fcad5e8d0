/*
 * custr.h — thin C-ABI of the B200-native string-column engine (libcustr.so).
 *
 * This is the drop-in boundary below the NVStrings / NVCategory / NVText classes: every entry point
 * replaces one reference method (cited per function as file:line under rapidsai/custrings) and takes only
 * plain pointers, sizes and opaque handles.  Conventions:
 *   - a column is an immutable Arrow-style triple resident in HBM: chars uint8[], offsets int32[n+1],
 *     validity bits (LSB-first, 1 = valid; NULL pointer = no nulls).  Replaces the reference's
 *     custring_view*[N] + object buffer (cpp/src/strings/NVStringsImpl.h:25-57).
 *   - functions returning int return >= 0 on success (usually the reference method's own return value) or a
 *     negative custr_status; custr_last_error() gives the message (thread local).  Nothing throws.
 *   - `devmem != 0` means the result / input array pointer is a device pointer, exactly like the
 *     reference's `bool devmem` / `bool todevice` arguments.
 *   - all work is enqueued on the stream set with custr_set_stream() (default: the legacy stream 0, as in
 *     the reference, cpp/src/strings/count.cu:66) and the call returns after the host-visible results exist.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with
 *     CUSTR_ERR_CUDA.
 */
#ifndef CUSTR_H
#define CUSTR_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct custr_column custr_column;      /* device string column (offsets, chars, validity) */
typedef struct custr_category custr_category;  /* dictionary: keys column + int32 values[n]       */

enum custr_status {
    CUSTR_OK = 0,
    CUSTR_ERR_ARG = -1,     /* reference returns -1 for null pattern/results (count.cu:61-62)             */
    CUSTR_ERR_INVALID = -2, /* reference throws std::invalid_argument (replace.cu:112, modify.cu:111)     */
    CUSTR_ERR_CUDA = -3,    /* reference throws std::runtime_error via CUDA_TRY (util.h:47-55)            */
    CUSTR_ERR_ALLOC = -4    /* reference throws std::runtime_error from device_alloc (util.inl:96-107)    */
};

const char* custr_last_error(void);
const char* custr_version(void);
int  custr_set_device(int device);
void custr_set_stream(void* cuda_stream);      /* cudaStream_t; thread local                              */
int  custr_sync(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
long long custr_launch_count(void);
/* Device memory is stream-ordered pool memory that the library keeps for reuse (the pool's release threshold is raised, and blocks of
 * 32 MiB and more sit in a per-thread cache of at most 6 GiB).  This returns what is not in use to the driver: for a process that
 * shares the GPU with another allocator and is done with a batch of calls. */
void custr_release_cached_memory(void);
/* name of the regex execution tier used by the last regex call on this thread ("bitstream", "pikevm", "bitcount",
 * "bitspans", "bitsplice", "chainspan", "literal") */
const char* custr_last_regex_tier(void);
/* force a tier for A/B testing: 0 = auto, 1 = exact Pike VM only, 2 = bitstream generic interpreter kernel (this also
 * switches the bit-stream forms of tokenize / split_record / literal replace / find's pre-filter off: per-row kernels),
 * 3 = boolean results from the window-at-a-time chain kernel (k_chain64) instead of the item-buffered one, and replace_re
 * through the per-row walk of the span streams instead of the streaming splice,
 * 4 = no ahead-of-time shape specialisation (run-time compiled plan kernel when available), 5 = neither (generic kernels) */
void custr_set_regex_tier(int tier);
/* Run-time compiled plan kernels: the chain kernel specialised for the pattern at hand with NVRTC (about 0.6 s on first use,
 * cached per process).  mode 0 = never, 1 = for columns of at least min_bytes chars whose plan no ahead-of-time shape covers
 * (default, 64 MiB), 2 = always.  min_bytes <= 0 keeps the current threshold.  Without libnvrtc the ahead-of-time kernels run. */
void custr_set_jit(int mode, long long min_bytes);
long long custr_jit_launch_count(void);   /* launches served by run-time compiled kernels so far */
const char* custr_jit_note(void);         /* why the last request on this thread was not served ("" if it was) */
/* tuning / A-B switch: fixed size of a work item of the bit-stream kernels in KiB of chars (4..32); anything else = the default,
 * chosen per column so that the item count is a whole number of rounds over the resident warps (at most 32 KiB) */
void custr_set_item_kib(int kib);
/* when on, regex calls bracket their dominant kernel(s) with CUDA events on the launch stream;
 * custr_last_kernel_ms() returns that device time for the last call on this thread (-1 if none) */
void custr_set_profiling(int on);
float custr_last_kernel_ms(void);

/* ---- column create / export  (NVStrings::create_from_offsets NVStrings.cu:109-119, create_from_array :74-86,
 *      create_offsets :402-482, to_host :266-346, set_null_bitarray :493-544, byte_count/len attrs.cu:32,72) ---- */
/* Copies host or device (devmem) buffers into a new column. validity may be NULL. Null rows are kept null
 * whatever their offsets say; bytes of null rows are not copied semantically (exported as zero length). */
custr_column* custr_create_from_offsets(const char* chars, int32_t count, const int32_t* offsets,
                                        const uint8_t* validity, int32_t nulls, int devmem);
/* Zero-copy view over caller-owned device buffers (must outlive the column). */
custr_column* custr_adopt_device(const char* chars, int32_t count, const int32_t* offsets, const uint8_t* validity, int32_t nulls);
/* Host array of NUL-terminated strings, NULL entries = null rows. */
custr_column* custr_create_from_array(const char* const* strs, uint32_t count);
/* NVStrings::create_from_index (NVStrings.h:98, NVStrings.cu:88-107): `pairs` = count {const char* ptr; size_t bytes;} entries
 * (the layout of std::pair<const char*,size_t>), in device memory when devmem != 0 else on the host; ptr always addresses
 * DEVICE memory, ptr == NULL is a null row.  stype = NVStrings::sorttype (0 none, 1 length, 2 name, 3 both). */
custr_column* custr_create_from_index(const void* pairs, uint32_t count, int devmem, int stype);
/* CUDA IPC (reference cpp/include/ipc_transfer.h, NVStrings::create_ipc_transfer / create_from_ipc): export packs the column
 * into one exportable device allocation (kept alive by `col`) and fills the handle; another process on the same GPU imports it
 * as a zero-copy column (the mapping closes when that column is freed).  The handle is plain bytes: ship it over any channel. */
typedef struct custr_ipc_handle {
    unsigned char handle[64];   /* cudaIpcMemHandle_t */
    int32_t n, nulls;
    int64_t chars_bytes;        /* chars at offset 0 */
    int64_t offsets_at;         /* byte offset of int32 offsets[n + 1] */
    int64_t validity_at;        /* byte offset of the validity bits, -1 if there are no nulls */
} custr_ipc_handle;
int custr_ipc_export(const custr_column* col, custr_ipc_handle* out);
custr_column* custr_ipc_import(const custr_ipc_handle* in);
void     custr_column_free(custr_column* col);
uint32_t custr_size(const custr_column* col);
int64_t  custr_chars_bytes(const custr_column* col);     /* bytes in the chars buffer                      */
int32_t  custr_null_count(const custr_column* col);
const char*    custr_chars_ptr(const custr_column* col);     /* device pointers                             */
const int32_t* custr_offsets_ptr(const custr_column* col);
const uint8_t* custr_validity_ptr(const custr_column* col);  /* NULL if no nulls                            */
/* Export to (chars, offsets[n+1], validity[(n+7)/8]); any of the three may be NULL to skip. Null rows export
 * as zero length (NVStrings.cu:427-432). Returns 0 like the reference (NVStrings.cu:402-482). */
int custr_create_offsets(const custr_column* col, char* chars, int32_t* offsets, uint8_t* validity, int devmem);
/* bitarray bit=1 valid; emptyIsNull also clears bits of empty rows. Returns the number of cleared bits. */
int custr_set_null_bitarray(const custr_column* col, uint8_t* bitarray, int empty_is_null, int devmem);
/* per-row byte length (-1 for null rows) / character length (-1 for null); either pointer may be NULL.
 * Returns total bytes (byte_count) or number of non-null rows (len). */
int64_t custr_byte_count(const custr_column* col, int32_t* lengths, int devmem);
int     custr_len(const custr_column* col, int32_t* lengths, int devmem);
/* MurmurHash3_x86_32 seed 31 over the row bytes, 0 for nulls (custring.inl:164-232, convert.cu:34-63). */
int custr_hash(const custr_column* col, uint32_t* results, int devmem);

/* ---- regex (Reprog NFA): NVStrings::contains_re count.cu:59-110, match :113-165, count_re :199-250 ---- */
int custr_contains_re(const custr_column* col, const char* pattern, uint8_t* results, int devmem);
int custr_match(const custr_column* col, const char* pattern, uint8_t* results, int devmem);
int custr_count_re(const custr_column* col, const char* pattern, int32_t* results, int devmem);
/* NVStrings::replace_re replace.cu:110-189; multi-pattern form replace_multi.cu:110-197 */
custr_column* custr_replace_re(const custr_column* col, const char* pattern, const char* repl, int32_t maxrepl);
custr_column* custr_replace_re_multi(const custr_column* col, const char* const* patterns, int32_t npatterns,
                                     const custr_column* repls);
/* NVStrings::replace_with_backrefs replace_backref.cu:122-213 (template parse regex/backref.h:31-57): every match is replaced
 * by `repl` with \1..\N substituted by the text of that capture group (\0 = whole match, unmatched / unknown group = nothing).
 * repl == NULL gives an all-null column; a NULL / empty pattern is INVALID. */
custr_column* custr_replace_with_backrefs(const custr_column* col, const char* pattern, const char* repl);
/* Capture-span callers ("next" rows of the scope table): NVStrings::findall findall.cu:99-170 (column c = c-th match of each
 * row, null where a row has fewer), findall_record findall_record.cu:97 (flat: all matches in row order + row offsets),
 * extract extract.cu:69-150 (one column per capture group of the FIRST match; null when the row does not match or the group
 * is empty).  Column-major variants write up to `cap` columns and return the number of columns. */
int custr_findall(const custr_column* col, const char* pattern, custr_column** out, int32_t cap);
int custr_findall_record(const custr_column* col, const char* pattern, custr_column** tokens, int32_t* row_offsets, int devmem);
int custr_extract(const custr_column* col, const char* pattern, custr_column** out, int32_t cap);
/* Debug/inspection: compile a pattern and print its instruction listing into buf (host). Returns #instructions. */
int custr_regex_describe(const char* pattern, char* buf, size_t buflen);

/* ---- literal find / replace: NVStrings::find find.cu:75-120, rfind :163-199, contains :237-272,
 *      startswith/endswith :316-387, find_multiple :202-233; replace modify.cu:109-192, multi :263-299 ---- */
int custr_find(const custr_column* col, const char* str, int32_t start, int32_t end, int32_t* results, int devmem);
int custr_rfind(const custr_column* col, const char* str, int32_t start, int32_t end, int32_t* results, int devmem);
int custr_contains(const custr_column* col, const char* str, uint8_t* results, int devmem);
int custr_startswith(const custr_column* col, const char* str, uint8_t* results, int devmem);
int custr_endswith(const custr_column* col, const char* str, uint8_t* results, int devmem);
/* NVStrings::find_from find.cu:123-160: per-row start / end character positions (NULL = 0 / end of string); the arrays live
 * where `devmem` says (the reference takes device pointers only).  NVStrings::match_strings find.cu:276-313: row-wise
 * equality with another column of the same size (two nulls are equal); returns the number of equal rows. */
int custr_find_from(const custr_column* col, const char* str, const int32_t* starts, const int32_t* ends, int32_t* results, int devmem);
int custr_match_strings(const custr_column* col, const custr_column* other, uint8_t* results, int devmem);
int custr_find_multiple(const custr_column* col, const custr_column* targets, int32_t* results, int devmem);
custr_column* custr_replace(const custr_column* col, const char* str, const char* repl, int32_t maxrepl);
custr_column* custr_replace_multi(const custr_column* col, const custr_column* targets, const custr_column* repls);

/* ---- split: NVStrings::split split.cu:734-822 (delimiter) / :863-956 (whitespace when delimiter==NULL),
 *      rsplit :960-1160, split_record :125-223 / :270-430 ---- */
/* Column-major. Writes up to `cap` new columns into out[]; returns the number of columns (call again with a
 * larger cap if the return value exceeds cap; extra columns are discarded). */
int custr_split(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** out, int32_t cap);
int custr_rsplit(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** out, int32_t cap);
/* Row-major, as ONE flat allocation instead of the reference's N objects: *tokens holds every token of every
 * row in row order, row_offsets[n+1] (caller array, host or device per devmem) delimits each row's slice;
 * null input rows get an empty slice and are distinguishable through the input validity. Returns total tokens. */
int custr_split_record(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** tokens,
                       int32_t* row_offsets, int devmem);
int custr_rsplit_record(const custr_column* col, const char* delimiter, int32_t maxsplit, custr_column** tokens,
                        int32_t* row_offsets, int devmem);
/* NVStrings::partition split.cu:1165-1262 / rpartition :1268-1372 (right != 0): three strings per row — [left, delimiter,
 * right] around the first / last delimiter, [row, "", ""] / ["", "", row] when it does not occur, three nulls for a null
 * row — returned row-major as ONE column of 3n rows instead of n objects.  NULL on error (empty delimiter: INVALID). */
custr_column* custr_partition(const custr_column* col, const char* delimiter, int right);
/* sub-range view [first, last) of a column as a new column (used to materialise one row of split_record). */
custr_column* custr_slice_rows(const custr_column* col, int32_t first, int32_t last);
/* gather rows by index (device or host int32 indices; negative/out-of-range -> null row). */
custr_column* custr_gather(const custr_column* col, const int32_t* indices, int32_t count, int devmem);

/* ---- cheap per-character attributes and transforms: strings/attrs.cu:115-445, case.cu:30-190,
 *      strip.cu:30-200, substr.cu:39-83 ---- */
/* kind: 0 isalnum, 1 isalpha, 2 isdigit, 3 isspace, 4 isdecimal, 5 isnumeric, 6 islower, 7 isupper, 8 is_empty; one bool per
 * row (null rows false; is_empty: true), returns the number of true rows */
int custr_is_class(const custr_column* col, int kind, uint8_t* results, int devmem);
custr_column* custr_case(const custr_column* col, int to_upper);                      /* lower() / upper()              */
custr_column* custr_strip(const custr_column* col, const char* to_strip, int side);   /* 0 strip, 1 lstrip, 2 rstrip; NULL = " \n\t" */
custr_column* custr_slice(const custr_column* col, int32_t start, int32_t stop, int32_t step);  /* characters [start, stop or end) */

/* ---- NVText::tokenize tokens.cu:123-155 (delimiter==NULL: whitespace), token_count :337-361 ---- */
custr_column* custr_tokenize(const custr_column* col, const char* delimiter);
int custr_token_count(const custr_column* col, const char* delimiter, uint32_t* results, int devmem);

/* ---- NVCategory: create_from_strings NVCategory.cu:327-356, NVCategoryImpl_init :220-304,
 *      get_keys :724-750, get_values :866-878, values_cptr :880-883 ---- */
custr_category* custr_category_create(const custr_column* const* cols, int32_t ncols);
void custr_category_free(custr_category* cat);
uint32_t custr_category_size(const custr_category* cat);
uint32_t custr_category_keys_size(const custr_category* cat);
custr_column* custr_category_keys(const custr_category* cat);              /* new column (copy)         */
int custr_category_values(const custr_category* cat, int32_t* results, int devmem);
const int32_t* custr_category_values_cptr(const custr_category* cat);      /* device pointer            */
/* Multi-GPU merge step (NVCategory::create_from_categories NVCategory.cu:430-514 is the reference's analogue):
 * given this shard's category and the union of all shards' keys (any order, duplicates allowed) returns a new
 * category whose keys are the sorted distinct union and whose values are remapped. */
custr_category* custr_category_remap_to_union(const custr_category* cat, const custr_column* all_keys);
/* Category algebra.  sorted != 0: NVCategory::create_from_categories / merge_and_remap NVCategory.cu:430-514,1339-1345 — keys =
 * sorted distinct union, values = every input's values remapped and appended in input order.  sorted == 0 (exactly two
 * inputs): NVCategory::merge_category :1223-1336 — keys = the first input's keys followed by the second's new keys, the first
 * values unchanged. */
custr_category* custr_category_merge(const custr_category* const* cats, int32_t ncats, int sorted);
/* Key-set algebra on an existing category (values keep their count, -1 = "no key"): op 0 = NVCategory::add_keys_and_remap
 * (NVCategory.cu:1375-1480: keys = sorted union with strs), 1 = remove_keys_and_remap (:1482-1565: values of removed keys become
 * -1), 2 = set_keys_and_remap (:1708-1820: keys = sorted distinct strs), 3 = remove_unused_keys_and_remap (:1567-1706, strs
 * ignored). */
custr_category* custr_category_keys_op(const custr_category* cat, const custr_column* strs, int op);
/* NVCategory::gather (:1142, remap == 0: same keys, values = pos, each in [0, keys) — the reference rejects -1 as well) / gather_and_remap (:1084, remap != 0:
 * keys = the keys pos uses, values remapped, each pos in [0, keys)); gather_strings (:1011): the key strings at pos[i].
 * pos: int32[count] on the device (devmem) or host.  Out-of-range positions: CUSTR_ERR_INVALID (reference: std::out_of_range). */
custr_category* custr_category_gather(const custr_category* cat, const int32_t* pos, int32_t count, int devmem, int remap);
custr_column* custr_category_gather_strings(const custr_category* cat, const int32_t* pos, int32_t count, int devmem);

/* ---- multi-GPU NVCategory: the one collective of the hot path.  One process per GPU; the column is sharded
 *      by contiguous row range.  custr_comm wraps an NCCL communicator (libnccl.so.2 bound at run time): rank 0 obtains a
 *      128-byte unique id, the launcher hands it to every rank (MPI / torch.distributed / a file), every rank creates its
 *      communicator.  custr_category_create_sharded = local NVCategory build (NVCategory.cu:327-356) -> ncclAllGather of
 *      every rank's distinct keys over NVLink -> global sorted key set + remap of the local values (create_from_categories
 *      math, NVCategory.cu:430-514).  Every rank returns the same keys; values index into them.  timings_ms: NULL or
 *      float[3] = {local build, key exchange, union + remap} in ms (CUDA events on the call's stream). ---- */
typedef struct custr_comm custr_comm;
int custr_comm_unique_id(void* id128);
custr_comm* custr_comm_create(int rank, int world, const void* id128);
void custr_comm_destroy(custr_comm* comm);
custr_category* custr_category_create_sharded(custr_comm* comm, const custr_column* local_rows, float* timings_ms);

#ifdef __cplusplus
}
#endif
#endif /* CUSTR_H */
