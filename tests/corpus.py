"""Shared test inputs: strings exercising empty / null / UTF-8 / newline / underscore cases and the pattern
families of the reference's own tests (cpp/tests/test_count.cu, python/tests/test_regex.py)."""
import random

STRINGS = [
    "abc de", "abcd", "xx héllo!", "", "a_b1 c", "abc_", "abc_ d", "_abcd", "ab_cd", "abcd_e", "a b c_de f", "ovér the", "1234",
    "00:0:00", None, "The quick brown fox", "jumps over the lazy dog", "hello\nworld", "\n", " ", "a", "ab\n", "tést", "ÀÉ",
    "12:34:56", "@home @work", "a.b", "AAA", "aXbXc", "line1\nline2\n", "0123456789", "foo(bar)", "[x]", "a|b", "日本語 テキスト",
    "emoji 😀 end", "tab\there", "x_y_z", "__init__", "CamelCaseWord", "5", "hello @abc @def world", "the", "ZZZ 123",
    "a" * 70, "word " * 30, "é" * 40, ("ab " * 50) + "abcd", "x" * 200 + " yyyy",
]

PATTERNS = [
    r"\b\w{4,}\b", r"\d+", r"@\w+", r"ov[eé]r", r"[tT]he", r"\d+:\d+", "a", "^a", "a$", r"\w+", r"\W+", r"\s", r"\S+", r"\D", r"[a-c]+",
    r"[^a-c]+", "a*", "a+b", "a?b", "(ab)+", "a|b", "ab|cd", r"(a|b)*c", r"\bfoo\b", r"\Bo", r"^$", r"^", r"$", r".", r".*", r".+x",
    r"a.c", r"l+", r"l{2}", r"l{2,}", r"l{1,2}o", r"(?:ab){2}", r"[0-9]{2}:", r"é", r"[à-ü]", r"\w*é", r"x*?y", r"a+?", r"a??b",
    r"a{2,3}?", r"\Aab", r"c\Z", r"^l", r"e$", r"o\n", r"\x41", r"\101x", r"\t", r"[\t ]", r"[\w]+@", r"[\W]", r"[\d_]+", r"[^\d]",
    r"日本", r"[一-龥]+", r"😀", r".😀", r"\b", r"\B", r"(a)(b)?", r"((a|b)c)*d", r"a||b", r"()", r"a+*", r"[a", r"a{", r"a{2", r"{2}",
    r"a)", r"(a", r"[]a]", r"[^]a]", r"[a-]", r"\\", r"a\\b", r"\.", r"[.]", r"\$", r"a{0}b", r"a{0,1}b", r"(ab){0,2}c", r"x{3}",
    r"[A-Z][a-z]+", r"\w+\s\w+", r"(\d+):(\d+)", r"e\b", r"\be", r"_\b", r"\b_", r"\w{4,}", r"\b\w+\b", r"[a-z]{3}\d", r"^\w+$",
    r"\w{2}\b", r"\bthe\b|\bfox\b", r"h.llo", r"^\s*$", r"\S\s\S", r"[^\n]+\n", r"\d{2}:\d{2}:\d{2}", r"(?:\w+ ){3}",
]

ALPHABET = list("abcxyz019_ .,:\n-@") + ["é", "ü", "日", "😀", "A", "Z"]
ATOMS = ["a", "b", "c", "x", "0", "1", "_", " ", "\\.", ".", "\\w", "\\W", "\\d", "\\D", "\\s", "\\S", "[abc]", "[^abc]", "[a-z]", "[0-9_]",
         "[^\\w]", "[\\d ]", "é", "[é-ü]", "日", "\\n", "[^\\n]", "😀"]


def random_strings(rng, count):
    out = []
    for _ in range(count):
        n = rng.choice([0, 1, 2, 3, 5, 8, 13, 21, 40, 90])
        out.append("".join(rng.choice(ALPHABET) for _ in range(n)))
    return out


def random_pattern(rng, depth=0):
    """Random pattern WITHOUT quantified groups (the reference spins forever on nullable loops such as (a*)*)."""
    def atom():
        r = rng.random()
        if depth < 2 and r < 0.15:
            return "(" + random_pattern(rng, depth + 1) + ")", True
        if depth < 2 and r < 0.25:
            return "(?:" + random_pattern(rng, depth + 1) + ")", True
        return rng.choice(ATOMS), False

    def piece():
        a, is_group = atom()
        r = rng.random()
        if is_group:
            return a + ("?" if r < 0.2 else "")
        if r < 0.12: a += "*"
        elif r < 0.24: a += "+"
        elif r < 0.34: a += "?"
        elif r < 0.38: a += "*?"
        elif r < 0.42: a += "+?"
        elif r < 0.45: a += "??"
        elif r < 0.52: a += "{%d}" % rng.randint(0, 3)
        elif r < 0.57: a += "{%d,}" % rng.randint(0, 2)
        elif r < 0.63:
            lo = rng.randint(0, 2)
            a += "{%d,%d}" % (lo, lo + rng.randint(0, 2))
        elif r < 0.65: a += "{1,2}?"
        return a

    def seq():
        parts = []
        for _ in range(rng.randint(1, 4)):
            r = rng.random()
            if r < 0.07: parts.append("\\b")
            elif r < 0.10: parts.append("\\B")
            elif r < 0.14: parts.append("^")
            elif r < 0.18: parts.append("$")
            else: parts.append(piece())
        return "".join(parts)

    return "|".join(seq() for _ in range(1 if rng.random() < 0.75 else rng.randint(2, 3)))


def random_patterns(seed, count):
    rng = random.Random(seed)
    return [random_pattern(rng) for _ in range(count)]
