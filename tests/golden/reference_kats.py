"""Known-answer vectors transcribed from the reference's OWN tests (inputs and expected literals only), each with
the file:line it comes from under /root/reference.  They pin the oracle (tests/test_oracle_pinned.py) and are
replayed through the C-ABI on the GPU (tests/test_gpu_kats.py)."""

COUNT_STRS = ["The quick brown @fox jumps", "ovér the", "lazy @dog", "1234", "00:0:00", None, ""]  # cpp/tests/test_count.cu:8-10
KATS = [
    # (op, strings, args, expected)                                              source
    ("contains", COUNT_STRS, ("é",), [False, True, False, False, False, False, False]),          # test_count.cu:18-21
    ("contains_re", COUNT_STRS, (r"\d+",), [False, False, False, True, True, False, False]),     # test_count.cu:24-27
    ("contains_re", COUNT_STRS, (r"@\w+",), [True, False, True, False, False, False, False]),    # test_count.cu:30-33
    ("match", COUNT_STRS, ("ov[eé]r",), [False, True, False, False, False, False, False]),       # test_count.cu:45-48
    ("match", COUNT_STRS, ("[tT]he",), [True, False, False, False, False, False, False]),        # test_count.cu:51-54
    ("match", COUNT_STRS, (r"\d+",), [False, False, False, True, True, False, False]),           # test_count.cu:57-60
    ("count_re", COUNT_STRS, ("[tT]he",), [1, 1, 0, 0, 0, 0, 0]),                                # test_count.cu:72-75
    ("count_re", COUNT_STRS, (r"@\w+",), [1, 0, 1, 0, 0, 0, 0]),                                 # test_count.cu:78-81
    ("count_re", COUNT_STRS, (r"\d+:\d+",), [0, 0, 0, 0, 1, 0, 0]),                              # test_count.cu:84-87
]

REPLACE_STRS = ["the quick brown fox jumps over the lazy dog", "the fat cat lays next to the other accénted cat",
                "a slow moving turtlé cannot catch the bird", "which can be composéd together to form a more complete",
                "thé result does not include the value in the sum in", "", "absent stop words"]  # cpp/tests/test_replace.cpp:7-12
KATS += [
    ("replace", REPLACE_STRS, ("the ", "++++ "),                                                 # test_replace.cpp:17-24
     ["++++ quick brown fox jumps over ++++ lazy dog", "++++ fat cat lays next to ++++ other accénted cat",
      "a slow moving turtlé cannot catch ++++ bird", "which can be composéd together to form a more complete",
      "thé result does not include ++++ value in ++++ sum in", "", "absent stop words"]),
    ("replace_re", REPLACE_STRS, (r"(\bin\b)|(\ba\b)|(\bthe\b)", "="),                           # test_replace.cpp:33-40
     ["= quick brown fox jumps over = lazy dog", "= fat cat lays next to = other accénted cat",
      "= slow moving turtlé cannot catch = bird", "which can be composéd together to form = more complete",
      "thé result does not include = value = = sum =", "", "absent stop words"]),
    ("replace_multi", REPLACE_STRS, (["the ", "a ", "to "], ["_ "]),                             # test_replace.cpp:49-60
     ["_ quick brown fox jumps over _ lazy dog", "_ fat cat lays next _ _ other accénted cat",
      "_ slow moving turtlé cannot catch _ bird", "which can be composéd together _ form _ more complete",
      "thé result does not include _ value in _ sum in", "", "absent stop words"]),
    ("replace_re_multi", REPLACE_STRS, ([r"\bthe\b", r"\ba\b", r"\bto\b"], ["", ".", "2"]),     # test_replace.cpp:71-81
     [" quick brown fox jumps over  lazy dog", " fat cat lays next 2  other accénted cat",
      ". slow moving turtlé cannot catch  bird", "which can be composéd together 2 form . more complete",
      "thé result does not include  value in  sum in", "", "absent stop words"]),
]

FIND_STRS = ["Héllo", "thesé", None, "ARE THE", "tést strings", ""]  # cpp/tests/test_find.cu:10
KATS += [
    ("find", FIND_STRS, ("é", 0, -1), [1, 4, -2, -1, 1, -1]),                                    # test_find.cu:33-35
    ("rfind", FIND_STRS, ("l", 0, -1), [3, -1, -2, -1, -1, -1]),                                 # test_find.cu:40-42
    ("contains", FIND_STRS, ("s",), [False, True, False, False, True, False]),                   # test_find.cu:69-71
    ("find_multiple", FIND_STRS, (["é", "e"],), [[1, -1], [4, 2], [-2, -2], [-1, -1], [1, -1], [-1, -1]]),  # test_find.cu:98-100
    ("endswith", FIND_STRS, ("E",), [False, False, False, True, False, False]),                  # test_find.cu:112-114
    ("startswith", FIND_STRS, ("t",), [False, True, False, False, True, False]),                 # test_find.cu:118-120
]

SPLIT_STRS = ["Héllo thesé", None, "are some", "tést String", ""]  # cpp/tests/test_split.cpp:6
KATS += [
    ("split", SPLIT_STRS, (None, -1), [["Héllo", None, "are", "tést", None], ["thesé", None, "some", "String", None]]),   # test_split.cpp:14-19
    ("split", SPLIT_STRS, ("s", -1), [["Héllo the", None, "are ", "té", ""], ["é", None, "ome", "t String", None]]),      # test_split.cpp:36-41
    ("split_record", SPLIT_STRS, (None, -1), [["Héllo", "thesé"], None, ["are", "some"], ["tést", "String"], [""]]),      # test_split.cpp:62-67
    ("split_record", SPLIT_STRS, ("s", -1), [["Héllo the", "é"], None, ["are ", "ome"], ["té", "t String"], [""]]),       # test_split.cpp:100-105
    ("rsplit", SPLIT_STRS, (None, -1), [["Héllo", None, "are", "tést", None], ["thesé", None, "some", "String", None]]),  # test_split.cpp:24-31
    ("rsplit", SPLIT_STRS, ("s", 2), [["Héllo the", None, "are ", "té", ""], ["é", None, "ome", "t String", None]]),      # test_split.cpp:46-53
    ("rsplit_record", SPLIT_STRS, (None, -1), [["Héllo", "thesé"], None, ["are", "some"], ["tést", "String"], [""]]),     # test_split.cpp:86-96
    ("partition", SPLIT_STRS, (" ",),                                                                                      # test_split.cpp:158-168
     [["Héllo", " ", "thesé"], [None, None, None], ["are", " ", "some"], ["tést", " ", "String"], ["", "", ""]]),
    ("rpartition", SPLIT_STRS, (" ",),                                                                                     # test_split.cpp:180-190
     [["Héllo", " ", "thesé"], [None, None, None], ["are", " ", "some"], ["tést", " ", "String"], ["", "", ""]]),
    ("replace_with_backrefs", REPLACE_STRS, (r"(\w) (\w)", r"\1-\2"),                                                      # test_replace.cpp:131-142
     ["the-quick-brown-fox-jumps-over-the-lazy-dog", "the-fat-cat-lays-next-to-the-other-accénted-cat",
      "a-slow-moving-turtlé-cannot-catch-the-bird", "which-can-be-composéd-together-to-form-a more-complete",
      "thé-result-does-not-include-the-value-in-the-sum-in", "", "absent-stop-words"]),
]

TEXT_STRS = ["the fox jumped over the dog", "the dog chased the cat", "the cat chased the mouse", None, "",
             "the mouse ate the cheese"]  # cpp/tests/test_text.cu:8-12
KATS += [
    ("tokenize", TEXT_STRS, (None,), ["the", "fox", "jumped", "over", "the", "dog", "the", "dog", "chased", "the", "cat", "the", "cat",
                                      "chased", "the", "mouse", "the", "mouse", "ate", "the", "cheese"]),               # test_text.cu:18-22
    ("token_count", TEXT_STRS, (" ",), [6, 5, 5, 0, 0, 5]),                                       # test_text.cu:33
]

HASH_STRS = ["thesé", None, "are", "the", "tést", "strings", ""]  # cpp/tests/test_convert.cu:9-10
KATS += [("hash", HASH_STRS, (), [126208335, 0, 3771471008, 2967174367, 1378466566, 3184694146, 1257683291])]  # test_convert.cu:17-18

# python/tests/test_category.py:19-61 and cpp/tests/cattest.cu
CAT_STRS = ["eee", "aaa", "eee", "ddd", "ccc", "ccc", "ccc", "eee", "aaa"]
KATS += [("category", CAT_STRS, (), (["aaa", "ccc", "ddd", "eee"], [3, 0, 3, 2, 1, 1, 1, 3, 0]))]

# oracle-generated micro-KATs for the headline pattern (the `_` cases differ from PCRE)
HEADLINE_STRS = ["abc de", "abcd", "xx héllo!", "", "a_b1 c", "abc_", "abc_ d", "_abcd", "ab_cd", "abcd_e", "a b c_de f", "ovér the", "1234",
                 "00:0:00", None]
KATS += [("contains_re", HEADLINE_STRS, (r"\b\w{4,}\b",),
          [False, True, True, False, True, False, False, True, True, True, True, True, True, False, False])]


def run_kat(op, col, args, api):
    """Evaluate one KAT with an adapter `api` exposing the ops over a column handle; returns a comparable value."""
    return getattr(api, op)(col, *args)
