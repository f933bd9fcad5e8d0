// C++ drop-in check: the reference's own gtest cases for the hot path (cpp/tests/test_count.cu, test_replace.cpp,
// test_find.cu, test_split.cpp, test_text.cu, test_convert.cu, cattest.cu) re-expressed against this repo's
// NVStrings / NVCategory / NVText classes with a tiny EXPECT shim (GoogleTest is not available offline).
// Host-memory result mode (devmem=false) throughout, like the Python bindings use.
#include <NVCategory.h>
#include <NVStrings.h>
#include <NVText.h>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

static int failures = 0;
#define EXPECT(cond)                                                            \
    do {                                                                        \
        if (!(cond)) { ++failures; printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); } \
    } while (0)

static std::vector<std::string> to_vec(NVStrings* s, std::vector<bool>* nulls = nullptr)
{
    unsigned n = s->size();
    std::vector<int> lens(n);
    s->byte_count(lens.data(), false);
    std::vector<std::vector<char>> bufs(n);
    std::vector<char*> ptrs(n);
    for (unsigned i = 0; i < n; ++i) { bufs[i].assign(lens[i] > 0 ? lens[i] + 1 : 1, 0); ptrs[i] = lens[i] >= 0 ? bufs[i].data() : nullptr; }
    s->to_host(ptrs.data(), 0, (int)n);
    std::vector<std::string> out;
    if (nulls) nulls->clear();
    for (unsigned i = 0; i < n; ++i) { out.push_back(ptrs[i] ? std::string(bufs[i].data()) : std::string()); if (nulls) nulls->push_back(lens[i] < 0); }
    return out;
}
static bool same(NVStrings* s, std::vector<const char*> want)
{
    std::vector<bool> nulls;
    std::vector<std::string> got = to_vec(s, &nulls);
    if (got.size() != want.size()) return false;
    for (size_t i = 0; i < want.size(); ++i) {
        if ((want[i] == nullptr) != nulls[i]) return false;
        if (want[i] && got[i] != want[i]) return false;
    }
    return true;
}

int main()
{
    {   // test_count.cu:8-100
        std::vector<const char*> h{"The quick brown @fox jumps", "ovér the", "lazy @dog", "1234", "00:0:00", nullptr, ""};
        NVStrings* s = NVStrings::create_from_array(h.data(), h.size());
        bool b[7];
        int rc = s->contains_re("\\d+", b, false);
        bool e1[] = {false, false, false, true, true, false, false};
        EXPECT(rc == 2 && !memcmp(b, e1, 7));
        s->contains_re("@\\w+", b, false);
        bool e2[] = {true, false, true, false, false, false, false};
        EXPECT(!memcmp(b, e2, 7));
        s->match("ov[eé]r", b, false);
        bool e3[] = {false, true, false, false, false, false, false};
        EXPECT(!memcmp(b, e3, 7));
        int c[7];
        s->count_re("\\d+:\\d+", c, false);
        int e4[] = {0, 0, 0, 0, 1, 0, 0};
        EXPECT(!memcmp(c, e4, sizeof(e4)));
        s->contains("é", b, false);
        EXPECT(!memcmp(b, e3, 7));
        EXPECT(s->contains_re(nullptr, b, false) == -1);
        NVStrings::destroy(s);
    }
    {   // test_replace.cpp:15-104
        std::vector<const char*> h{"the quick brown fox jumps over the lazy dog", "the fat cat lays next to the other accénted cat",
                                   "a slow moving turtlé cannot catch the bird", "which can be composéd together to form a more complete",
                                   "thé result does not include the value in the sum in", "", "absent stop words"};
        NVStrings* s = NVStrings::create_from_array(h.data(), h.size());
        NVStrings* g = s->replace("the ", "++++ ");
        EXPECT(same(g, {"++++ quick brown fox jumps over ++++ lazy dog", "++++ fat cat lays next to ++++ other accénted cat",
                        "a slow moving turtlé cannot catch ++++ bird", "which can be composéd together to form a more complete",
                        "thé result does not include ++++ value in ++++ sum in", "", "absent stop words"}));
        NVStrings::destroy(g);
        g = s->replace_re("(\\bin\\b)|(\\ba\\b)|(\\bthe\\b)", "=");
        EXPECT(same(g, {"= quick brown fox jumps over = lazy dog", "= fat cat lays next to = other accénted cat",
                        "= slow moving turtlé cannot catch = bird", "which can be composéd together to form = more complete",
                        "thé result does not include = value = = sum =", "", "absent stop words"}));
        NVStrings::destroy(g);
        std::vector<const char*> pats{"\\bthe\\b", "\\ba\\b", "\\bto\\b"}, rp{"", ".", "2"};
        NVStrings* r = NVStrings::create_from_array(rp.data(), rp.size());
        g = s->replace_re(pats, *r);
        EXPECT(same(g, {" quick brown fox jumps over  lazy dog", " fat cat lays next 2  other accénted cat",
                        ". slow moving turtlé cannot catch  bird", "which can be composéd together 2 form . more complete",
                        "thé result does not include  value in  sum in", "", "absent stop words"}));
        NVStrings::destroy(g);
        NVStrings::destroy(r);
        bool threw = false;
        try { s->replace_re("", "x"); } catch (const std::invalid_argument&) { threw = true; }
        EXPECT(threw);
        threw = false;
        try { s->replace("", "x"); } catch (const std::invalid_argument&) { threw = true; }
        EXPECT(threw);
        NVStrings::destroy(s);
    }
    {   // test_find.cu:10-131
        std::vector<const char*> h{"Héllo", "thesé", nullptr, "ARE THE", "tést strings", ""};
        NVStrings* s = NVStrings::create_from_array(h.data(), h.size());
        int r[6];
        s->find("é", 0, -1, r, false);
        int e1[] = {1, 4, -2, -1, 1, -1};
        EXPECT(!memcmp(r, e1, sizeof(e1)));
        s->rfind("l", 0, -1, r, false);
        int e2[] = {3, -1, -2, -1, -1, -1};
        EXPECT(!memcmp(r, e2, sizeof(e2)));
        bool b[6];
        s->endswith("E", b, false);
        bool e3[] = {false, false, false, true, false, false};
        EXPECT(!memcmp(b, e3, 6));
        NVStrings::destroy(s);
    }
    {   // test_split.cpp:10-205
        std::vector<const char*> h{"Héllo thesé", nullptr, "are some", "tést String", ""};
        NVStrings* s = NVStrings::create_from_array(h.data(), h.size());
        std::vector<NVStrings*> cols;
        EXPECT(s->split(-1, cols) == 2);
        EXPECT(same(cols[0], {"Héllo", nullptr, "are", "tést", nullptr}) && same(cols[1], {"thesé", nullptr, "some", "String", nullptr}));
        for (auto c : cols) NVStrings::destroy(c);
        cols.clear();
        EXPECT(s->split("s", -1, cols) == 2);
        EXPECT(same(cols[0], {"Héllo the", nullptr, "are ", "té", ""}) && same(cols[1], {"é", nullptr, "ome", "t String", nullptr}));
        for (auto c : cols) NVStrings::destroy(c);
        std::vector<NVStrings*> rows;
        EXPECT(s->split_record("s", -1, rows) == 7);
        EXPECT(rows.size() == 5 && rows[1] == nullptr && same(rows[0], {"Héllo the", "é"}) && same(rows[3], {"té", "t String"}) && same(rows[4], {""}));
        for (auto r : rows) if (r) NVStrings::destroy(r);
        // right-to-left and partition forms: test_split.cpp:24-56,86-110,158-205
        cols.clear();
        EXPECT(s->rsplit(-1, cols) == 2);
        EXPECT(same(cols[0], {"Héllo", nullptr, "are", "tést", nullptr}) && same(cols[1], {"thesé", nullptr, "some", "String", nullptr}));
        for (auto c : cols) NVStrings::destroy(c);
        cols.clear();
        EXPECT(s->rsplit("s", 2, cols) == 2);
        EXPECT(same(cols[0], {"Héllo the", nullptr, "are ", "té", ""}) && same(cols[1], {"é", nullptr, "ome", "t String", nullptr}));
        for (auto c : cols) NVStrings::destroy(c);
        rows.clear();
        s->rsplit_record(-1, rows);
        EXPECT(rows.size() == 5 && rows[1] == nullptr && same(rows[0], {"Héllo", "thesé"}) && same(rows[2], {"are", "some"}) && same(rows[4], {""}));
        for (auto r : rows) if (r) NVStrings::destroy(r);
        rows.clear();
        EXPECT(s->partition(" ", rows) == 5);
        EXPECT(same(rows[0], {"Héllo", " ", "thesé"}) && same(rows[1], {nullptr, nullptr, nullptr}) && same(rows[3], {"tést", " ", "String"}) &&
               same(rows[4], {"", "", ""}));
        for (auto r : rows) if (r) NVStrings::destroy(r);
        rows.clear();
        EXPECT(s->rpartition(" ", rows) == 5);
        EXPECT(same(rows[0], {"Héllo", " ", "thesé"}) && same(rows[1], {nullptr, nullptr, nullptr}) && same(rows[2], {"are", " ", "some"}));
        for (auto r : rows) if (r) NVStrings::destroy(r);
        NVStrings::destroy(s);
    }
    {   // test_replace.cpp:131-148
        std::vector<const char*> h{"the quick brown fox jumps over the lazy dog", "the fat cat lays next to the other accénted cat",
                                   "a slow moving turtlé cannot catch the bird", "which can be composéd together to form a more complete",
                                   "thé result does not include the value in the sum in", "", "absent stop words"};
        NVStrings* s = NVStrings::create_from_array(h.data(), h.size());
        NVStrings* got = s->replace_with_backrefs("(\\w) (\\w)", "\\1-\\2");
        EXPECT(same(got, {"the-quick-brown-fox-jumps-over-the-lazy-dog", "the-fat-cat-lays-next-to-the-other-accénted-cat",
                          "a-slow-moving-turtlé-cannot-catch-the-bird", "which-can-be-composéd-together-to-form-a more-complete",
                          "thé-result-does-not-include-the-value-in-the-sum-in", "", "absent-stop-words"}));
        NVStrings::destroy(got);
        NVStrings::destroy(s);
    }
    {   // test_text.cu:15-39, test_convert.cu:9-23, cattest.cu
        std::vector<const char*> t{"the fox jumped over the dog", "the dog chased the cat", nullptr, "", "the mouse ate the cheese"};
        NVStrings* s = NVStrings::create_from_array(t.data(), t.size());
        NVStrings* tok = NVText::tokenize(*s);
        EXPECT(tok->size() == 16);
        NVStrings::destroy(tok);
        unsigned cnt[5];
        NVText::token_count(*s, " ", cnt, false);
        EXPECT(cnt[0] == 6 && cnt[1] == 5 && cnt[2] == 0 && cnt[3] == 0 && cnt[4] == 5);
        NVStrings::destroy(s);
        std::vector<const char*> hs{"thesé", nullptr, "are", "the", "tést", "strings", ""};
        s = NVStrings::create_from_array(hs.data(), hs.size());
        unsigned hv[7];
        s->hash(hv, false);
        unsigned want[] = {126208335u, 0u, 3771471008u, 2967174367u, 1378466566u, 3184694146u, 1257683291u};
        EXPECT(!memcmp(hv, want, sizeof(want)));
        NVStrings::destroy(s);
        std::vector<const char*> cs{"eee", "aaa", "eee", "ddd", "ccc", "ccc", "ccc", "eee", "aaa"};
        NVCategory* cat = NVCategory::create_from_array(cs.data(), cs.size());
        int vals[9];
        cat->get_values(vals, false);
        int wv[] = {3, 0, 3, 2, 1, 1, 1, 3, 0};
        EXPECT(!memcmp(vals, wv, sizeof(wv)) && cat->keys_size() == 4);
        NVStrings* keys = cat->get_keys();
        EXPECT(same(keys, {"aaa", "ccc", "ddd", "eee"}));
        NVStrings::destroy(keys);
        NVCategory::destroy(cat);
    }
    printf(failures ? "FAILED %d checks\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
