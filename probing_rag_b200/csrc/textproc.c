/* Host-side tokenizer of the corpus front-end (libprtext.so, plain C, no CUDA).
 *
 * What llama-index's BM25Retriever does to every passage before indexing (/root/reference/exp_rag.py:242 ->
 * bm25s.tokenize): lower-case, re.findall(r"(?u)\b\w\w+\b"), i.e. maximal runs of Unicode
 * alphanumerics / '_' of at least two code points.  The Python restatement (probing_rag_b200/text.py) spends ~45 us
 * per passage in re + dict look-ups; this does the same at memory speed and hands back dense SURFACE-TOKEN ids
 * (first-seen order) that text.py maps to stems with numpy.
 *
 * Exactness: the word-character bitmap and the lower-case map are GENERATED from the running Python interpreter
 * (unicode_tables.h, written by probing_rag_b200/build.py), for the Basic Multilingual Plane.  A document that
 * contains anything the tables cannot express exactly -- a code point above U+FFFF, a character whose lower() is not
 * one BMP code point (U+0130), the context-sensitive capital sigma (U+03A3), a token longer than the buffer -- is
 * not tokenized here: pt_encode stops in front of it and the caller runs the Python path for that one document
 * (interning its tokens through pt_intern_many, so ids stay in first-seen order).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "unicode_tables.h" /* pt_isword[8192], pt_lower[65536] */

#define PT_MAX_TOKEN 1024

#define PT_INLINE 16
typedef struct {
    uint64_t hash;
    uint32_t len;
    int32_t id;                    /* -1 = empty slot */
    uint64_t head[2];              /* the first PT_INLINE bytes of the token, zero padded: most look-ups never touch the arena */
} pt_slot;                         /* 32 bytes: two per cache line, one miss per probe */

typedef struct pt_table {
    pt_slot *slots;
    uint64_t cap, n; /* cap is a power of two */
    char *arena;
    uint64_t arena_len, arena_cap;
    uint64_t *tok_off; /* arena offset of token id i; tok_off[n] = arena_len */
    uint64_t tok_cap;
} pt_table;

/* the (up to) eight bytes at p as a little-endian word, bytes beyond n zeroed.  ALWAYS reads eight bytes: every
 * caller hands in a buffer padded by PT_PAD (fixed-size loads instead of variable-length memcpy/memcmp calls are what
 * makes a look-up cheap) */
#define PT_PAD 16
static inline uint64_t load_masked(const unsigned char *p, uint32_t n)
{
    uint64_t v;
    memcpy(&v, p, 8);
    return n >= 8 ? v : (n ? (v & ((1ull << (8 * n)) - 1)) : 0);
}

static inline uint64_t hash_bytes(const unsigned char *p, uint32_t n)
{
    uint64_t h = 0x9e3779b97f4a7c15ull ^ n;
    for (uint32_t i = 0; i < n; i += 8) {
        h = (h ^ load_masked(p + i, n - i)) * 0xff51afd7ed558ccdull;
        h ^= h >> 32;
    }
    return h;
}

pt_table *pt_create(void)
{
    pt_table *t = (pt_table *)calloc(1, sizeof(pt_table));
    if (!t) return NULL;
    t->cap = 1u << 16;
    t->slots = (pt_slot *)malloc(t->cap * sizeof(pt_slot));
    t->arena_cap = 1u << 20;
    t->arena = (char *)malloc(t->arena_cap);
    t->tok_cap = 1u << 16;
    t->tok_off = (uint64_t *)malloc((t->tok_cap + 1) * sizeof(uint64_t));
    if (!t->slots || !t->arena || !t->tok_off) return NULL;
    for (uint64_t i = 0; i < t->cap; ++i) t->slots[i].id = -1;
    t->tok_off[0] = 0;
    return t;
}

void pt_destroy(pt_table *t)
{
    if (!t) return;
    free(t->slots);
    free(t->arena);
    free(t->tok_off);
    free(t);
}

const char *pt_unidata_version(void) { return PT_UNIDATA_VERSION; }

int64_t pt_size(const pt_table *t) { return t ? (int64_t)t->n : 0; }

static int grow(pt_table *t)
{
    const uint64_t ncap = t->cap * 2;
    pt_slot *ns = (pt_slot *)malloc(ncap * sizeof(pt_slot));
    if (!ns) return -1;
    for (uint64_t i = 0; i < ncap; ++i) ns[i].id = -1;
    for (uint64_t i = 0; i < t->cap; ++i) {
        if (t->slots[i].id < 0) continue;
        uint64_t j = t->slots[i].hash & (ncap - 1);
        while (ns[j].id >= 0) j = (j + 1) & (ncap - 1);
        ns[j] = t->slots[i];
    }
    free(t->slots);
    t->slots = ns;
    t->cap = ncap;
    return 0;
}

/* id of the token (bytes p[0..n), in a buffer readable PT_PAD bytes past the end), added if new; -1 = out of memory */
static int32_t intern(pt_table *t, const unsigned char *p, uint32_t n)
{
    const uint64_t h = hash_bytes(p, n);
    const uint64_t k0 = load_masked(p, n), k1 = n > 8 ? load_masked(p + 8, n - 8) : 0;
    uint64_t j = h & (t->cap - 1);
    while (t->slots[j].id >= 0) {
        const pt_slot *s = &t->slots[j];
        if (s->hash == h && s->len == n && s->head[0] == k0 && s->head[1] == k1 &&
            (n <= PT_INLINE || memcmp(t->arena + t->tok_off[s->id] + PT_INLINE, p + PT_INLINE, n - PT_INLINE) == 0))
            return s->id;
        j = (j + 1) & (t->cap - 1);
    }
    if (t->n >= 0x7ffffff0ull) return -1;
    if (t->arena_len + n > t->arena_cap) {
        uint64_t nc = t->arena_cap * 2;
        while (nc < t->arena_len + n) nc *= 2;
        char *na = (char *)realloc(t->arena, nc);
        if (!na) return -1;
        t->arena = na;
        t->arena_cap = nc;
    }
    if (t->n + 1 > t->tok_cap) {
        uint64_t *no = (uint64_t *)realloc(t->tok_off, (t->tok_cap * 2 + 1) * sizeof(uint64_t));
        if (!no) return -1;
        t->tok_off = no;
        t->tok_cap *= 2;
    }
    memcpy(t->arena + t->arena_len, p, n);
    pt_slot *s = &t->slots[j];
    s->hash = h;
    s->len = n;
    s->head[0] = k0;
    s->head[1] = k1;
    s->id = (int32_t)t->n;
    t->arena_len += n;
    t->n += 1;
    t->tok_off[t->n] = t->arena_len;
    const int32_t id = s->id;
    if (t->n * 2 > t->cap && grow(t) != 0) return -1;
    return id;
}

static inline int isword(uint32_t cp) { return (pt_isword[cp >> 3] >> (cp & 7)) & 1; }

/* Tokenize documents [first, n_docs) of the UTF-8 buffer (document d = buf[offs[d] .. offs[d+1])) until one needs the
 * Python path.  Appends surface-token ids to out_ids starting at *n_tok (capacity out_cap), writes counts[d].
 * Returns the index of the first document NOT processed (n_docs when all were), or -1 on out-of-memory / capacity. */
int64_t pt_encode(pt_table *t, const unsigned char *buf, const int64_t *offs, int64_t first, int64_t n_docs, int32_t *out_ids,
                  int64_t out_cap, int64_t *n_tok, int32_t *counts)
{
    unsigned char tok[PT_MAX_TOKEN + 4 + PT_PAD];
    unsigned char ascii[128]; /* lower-cased byte of a word character, 0 for a separator */
    for (uint32_t c = 0; c < 128; ++c) ascii[c] = (pt_lower[c] < 128 && isword(pt_lower[c])) ? (unsigned char)pt_lower[c] : 0;
    int64_t nt = *n_tok;
    for (int64_t d = first; d < n_docs; ++d) {
        const unsigned char *p = buf + offs[d], *e = buf + offs[d + 1];
        const int64_t tok0 = nt;
        uint32_t len = 0, ncp = 0;
        int fallback = 0;
#define PT_FLUSH()                                      \
    do {                                                \
        if (ncp >= 2) {                                 \
            if (nt >= out_cap) return -1;               \
            const int32_t id_ = intern(t, tok, len);    \
            if (id_ < 0) return -1;                     \
            out_ids[nt++] = id_;                        \
        }                                               \
        len = 0;                                        \
        ncp = 0;                                        \
    } while (0)
        while (p < e) {
            const unsigned char c = *p;
            if (c < 0x80) { /* the common case: one table look-up per byte */
                const unsigned char lo = ascii[c];
                ++p;
                if (lo) {
                    if (len + 3 > PT_MAX_TOKEN) {
                        fallback = 1;
                        break;
                    }
                    tok[len++] = lo;
                    ++ncp;
                } else if (len) {
                    PT_FLUSH();
                }
                continue;
            }
            uint32_t cp;
            if (c < 0xe0) {
                cp = ((uint32_t)(c & 0x1f) << 6) | (p[1] & 0x3f);
                p += 2;
            } else if (c < 0xf0) {
                cp = ((uint32_t)(c & 0x0f) << 12) | ((uint32_t)(p[1] & 0x3f) << 6) | (p[2] & 0x3f);
                p += 3;
            } else {
                fallback = 1; /* beyond the BMP */
                break;
            }
            const uint32_t lo = pt_lower[cp];
            if (lo == 0xffffu || len + 3 > PT_MAX_TOKEN) {
                fallback = 1;
                break;
            }
            if (isword(lo)) {
                if (lo < 0x80) {
                    tok[len++] = (unsigned char)lo;
                } else if (lo < 0x800) {
                    tok[len++] = (unsigned char)(0xc0 | (lo >> 6));
                    tok[len++] = (unsigned char)(0x80 | (lo & 0x3f));
                } else {
                    tok[len++] = (unsigned char)(0xe0 | (lo >> 12));
                    tok[len++] = (unsigned char)(0x80 | ((lo >> 6) & 0x3f));
                    tok[len++] = (unsigned char)(0x80 | (lo & 0x3f));
                }
                ++ncp;
            } else {
                PT_FLUSH();
            }
        }
        if (fallback) { /* the tokens interned so far stay in the table: the caller re-interns them in the same order */
            *n_tok = tok0;
            return d;
        }
        PT_FLUSH();
#undef PT_FLUSH
        counts[d] = (int32_t)(nt - tok0);
        *n_tok = nt;
    }
    return n_docs;
}

/* ids of n already lower-cased, already split tokens (token i = buf[offs[i] .. offs[i+1])); 0, or -1 on out-of-memory */
int pt_intern_many(pt_table *t, const unsigned char *buf, const int64_t *offs, int64_t n, int32_t *out_ids)
{
    unsigned char local[4096 + PT_PAD];
    for (int64_t i = 0; i < n; ++i) {
        const uint32_t len = (uint32_t)(offs[i + 1] - offs[i]);
        unsigned char *tmp = len <= 4096 ? local : (unsigned char *)malloc((size_t)len + PT_PAD); /* padded copy, see load_masked */
        if (!tmp) return -1;
        memcpy(tmp, buf + offs[i], len);
        const int32_t id = intern(t, tmp, len);
        if (tmp != local) free(tmp);
        if (id < 0) return -1;
        out_ids[i] = id;
    }
    return 0;
}

/* bytes and offsets of the tokens with ids [from, size): out_offs gets size - from + 1 offsets relative to the first
 * byte copied; returns the number of bytes (call with out_buf = NULL to size the buffer) */
int64_t pt_tokens_since(const pt_table *t, int64_t from, unsigned char *out_buf, int64_t out_cap, int64_t *out_offs)
{
    if (from < 0 || (uint64_t)from > t->n) return -1;
    const uint64_t b0 = t->tok_off[from], b1 = t->arena_len;
    if (!out_buf) return (int64_t)(b1 - b0);
    if ((int64_t)(b1 - b0) > out_cap) return -1;
    memcpy(out_buf, t->arena + b0, b1 - b0);
    for (uint64_t i = (uint64_t)from; i <= t->n; ++i) out_offs[i - from] = (int64_t)(t->tok_off[i] - b0);
    return (int64_t)(b1 - b0);
}

/* ------------------------------------------------------------------------------------------------------------------
 * Snowball English ("Porter2") stemming of ASCII words -- the algorithm PyStemmer's Stemmer("english") implements
 * (what llama-index's BM25Retriever stems with), restated from its published definition exactly as
 * probing_rag_b200/text.py:porter2_stem does (the two are compared word for word in tests/test_text.py).  Words with
 * a byte >= 0x80 are the caller's business: the algorithm counts letters, not bytes.
 * ------------------------------------------------------------------------------------------------------------------ */
static int st_v(unsigned char c) { return c == 'a' || c == 'e' || c == 'i' || c == 'o' || c == 'u' || c == 'y'; }

static int st_ends(const unsigned char *w, int n, const char *suf)
{
    const int m = (int)strlen(suf);
    return n >= m && memcmp(w + n - m, suf, (size_t)m) == 0;
}

static int st_is(const unsigned char *w, int n, const char *s) { return (int)strlen(s) == n && memcmp(w, s, (size_t)n) == 0; }

static int st_has_vowel(const unsigned char *w, int n)
{
    for (int i = 0; i < n; ++i)
        if (st_v(w[i])) return 1;
    return 0;
}

static int st_region(const unsigned char *w, int n, int start)
{
    for (int i = start + 1; i < n; ++i)
        if (!st_v(w[i]) && st_v(w[i - 1])) return i + 1;
    return n;
}

static int st_short_syllable(const unsigned char *w, int n)
{
    if (n == 2) return st_v(w[0]) && !st_v(w[1]);
    if (n >= 3)
        return !st_v(w[n - 3]) && st_v(w[n - 2]) && !st_v(w[n - 1]) && w[n - 1] != 'w' && w[n - 1] != 'x' && w[n - 1] != 'Y';
    return 0;
}

static int st_finish(unsigned char *w, int n)
{
    for (int i = 0; i < n; ++i)
        if (w[i] == 'Y') w[i] = 'y';
    return n;
}

/* stems word[0..n) into out (room for n + 2 bytes); returns the stem's length */
static int st_stem(const unsigned char *word, int n, unsigned char *w)
{
    static const char *const exc[][2] = {
        {"skis", "ski"},     {"skies", "sky"},   {"dying", "die"},  {"lying", "lie"},   {"tying", "tie"},   {"idly", "idl"},
        {"gently", "gentl"}, {"ugly", "ugli"},   {"early", "earli"}, {"only", "onli"},   {"singly", "singl"}, {"sky", "sky"},
        {"news", "news"},    {"howe", "howe"},   {"atlas", "atlas"}, {"cosmos", "cosmos"}, {"bias", "bias"},  {"andes", "andes"}};
    static const char *const exc1a[] = {"inning", "outing", "canning", "herring", "earring", "proceed", "exceed", "succeed"};
    static const char *const step2[][2] = {
        {"ization", "ize"}, {"ational", "ate"}, {"fulness", "ful"}, {"ousness", "ous"}, {"iveness", "ive"}, {"tional", "tion"},
        {"biliti", "ble"},  {"lessli", "less"}, {"entli", "ent"},   {"ation", "ate"},   {"alism", "al"},    {"aliti", "al"},
        {"ousli", "ous"},   {"iviti", "ive"},   {"fulli", "ful"},   {"enci", "ence"},   {"anci", "ance"},   {"abli", "able"},
        {"izer", "ize"},    {"ator", "ate"},    {"alli", "al"},     {"bli", "ble"},     {"ogi", "og"},      {"li", ""}};
    static const char *const step3[][2] = {{"ational", "ate"}, {"tional", "tion"}, {"alize", "al"}, {"icate", "ic"}, {"iciti", "ic"},
                                           {"ative", ""},      {"ical", "ic"},     {"ness", ""},    {"ful", ""}};
    static const char *const step4[] = {"ement", "ance", "ence", "able", "ible", "ment", "ant", "ent", "ism",
                                        "ate",   "iti",  "ous",  "ive",  "ize",  "ion",  "al",  "er",  "ic"};
    static const char *const doubles[] = {"bb", "dd", "ff", "gg", "mm", "nn", "pp", "rr", "tt"};

    if (n <= 2) {
        memcpy(w, word, (size_t)n);
        return n;
    }
    for (unsigned i = 0; i < sizeof(exc) / sizeof(exc[0]); ++i)
        if (st_is(word, n, exc[i][0])) {
            const int m = (int)strlen(exc[i][1]);
            memcpy(w, exc[i][1], (size_t)m);
            return m;
        }
    if (word[0] == '\'') {
        ++word;
        --n;
    }
    memcpy(w, word, (size_t)n);
    if (n <= 2) return n;
    if (w[0] == 'y') w[0] = 'Y';
    for (int i = 1; i < n; ++i)
        if (w[i] == 'y' && st_v(w[i - 1])) w[i] = 'Y';
    int r1;
    if (n >= 5 && (memcmp(w, "gener", 5) == 0 || memcmp(w, "arsen", 5) == 0))
        r1 = 5;
    else if (n >= 6 && memcmp(w, "commun", 6) == 0)
        r1 = 6;
    else
        r1 = st_region(w, n, 0);
    const int r2 = st_region(w, n, r1);

    /* step 0 */
    if (st_ends(w, n, "'s'"))
        n -= 3;
    else if (st_ends(w, n, "'s"))
        n -= 2;
    else if (st_ends(w, n, "'"))
        n -= 1;
    /* step 1a */
    if (st_ends(w, n, "sses"))
        n -= 2;
    else if (st_ends(w, n, "ied") || st_ends(w, n, "ies"))
        n -= (n > 4) ? 2 : 1;
    else if (st_ends(w, n, "us") || st_ends(w, n, "ss"))
        ;
    else if (st_ends(w, n, "s")) {
        if (n >= 2 && st_has_vowel(w, n - 2)) n -= 1;
    }
    for (unsigned i = 0; i < sizeof(exc1a) / sizeof(exc1a[0]); ++i)
        if (st_is(w, n, exc1a[i])) return st_finish(w, n);
    /* step 1b */
    if (st_ends(w, n, "eedly")) {
        if (n - 5 >= r1) n -= 3;
    } else if (st_ends(w, n, "eed")) {
        if (n - 3 >= r1) n -= 1;
    } else {
        static const char *const sufs[] = {"ingly", "edly", "ing", "ed"};
        for (int k = 0; k < 4; ++k) {
            if (!st_ends(w, n, sufs[k])) continue;
            const int m = n - (int)strlen(sufs[k]);
            if (st_has_vowel(w, m)) {
                n = m;
                int dbl = 0;
                for (int q = 0; q < 9; ++q) dbl |= st_ends(w, n, doubles[q]);
                if (st_ends(w, n, "at") || st_ends(w, n, "bl") || st_ends(w, n, "iz"))
                    w[n++] = 'e';
                else if (dbl)
                    n -= 1;
                else if (st_short_syllable(w, n) && r1 >= n)
                    w[n++] = 'e';
            }
            break;
        }
    }
    /* step 1c */
    if (n > 2 && (w[n - 1] == 'y' || w[n - 1] == 'Y') && !st_v(w[n - 2])) w[n - 1] = 'i';
    /* step 2 */
    for (unsigned i = 0; i < sizeof(step2) / sizeof(step2[0]); ++i) {
        const char *suf = step2[i][0], *rep = step2[i][1];
        if (!st_ends(w, n, suf)) continue;
        const int ls = (int)strlen(suf), lr = (int)strlen(rep);
        if (n - ls >= r1) {
            if (strcmp(suf, "ogi") == 0) {
                if (n - 3 >= 1 && w[n - 4] == 'l') {
                    memcpy(w + n - 3, rep, (size_t)lr);
                    n += lr - 3;
                }
            } else if (strcmp(suf, "li") == 0) {
                if (n > 2 && strchr("cdeghkmnrt", w[n - 3]) != NULL && w[n - 3] != 0) n -= 2;
            } else {
                memcpy(w + n - ls, rep, (size_t)lr);
                n += lr - ls;
            }
        }
        break;
    }
    /* step 3 */
    for (unsigned i = 0; i < sizeof(step3) / sizeof(step3[0]); ++i) {
        const char *suf = step3[i][0], *rep = step3[i][1];
        if (!st_ends(w, n, suf)) continue;
        const int ls = (int)strlen(suf), lr = (int)strlen(rep);
        if (n - ls >= r1) {
            if (strcmp(suf, "ative") == 0) {
                if (n - 5 >= r2) n -= 5;
            } else {
                memcpy(w + n - ls, rep, (size_t)lr);
                n += lr - ls;
            }
        }
        break;
    }
    /* step 4 */
    for (unsigned i = 0; i < sizeof(step4) / sizeof(step4[0]); ++i) {
        const char *suf = step4[i];
        if (!st_ends(w, n, suf)) continue;
        const int ls = (int)strlen(suf);
        if (n - ls >= r2) {
            if (strcmp(suf, "ion") == 0) {
                if (n > 3 && (w[n - 4] == 's' || w[n - 4] == 't')) n -= 3;
            } else {
                n -= ls;
            }
        }
        break;
    }
    /* step 5 */
    if (st_ends(w, n, "e")) {
        if (n - 1 >= r2 || (n - 1 >= r1 && !st_short_syllable(w, n - 1))) n -= 1;
    } else if (st_ends(w, n, "l")) {
        if (n - 1 >= r2 && n > 1 && w[n - 2] == 'l') n -= 1;
    }
    return st_finish(w, n);
}

/* stems n ASCII words (word i = buf[offs[i] .. offs[i+1])) into out (capacity >= total bytes + 2 n), out_offs[n + 1] */
void pt_stem_many(const unsigned char *buf, const int64_t *offs, int64_t n, unsigned char *out, int64_t *out_offs)
{
    int64_t o = 0;
    out_offs[0] = 0;
    for (int64_t i = 0; i < n; ++i) {
        o += st_stem(buf + offs[i], (int)(offs[i + 1] - offs[i]), out + o);
        out_offs[i + 1] = o;
    }
}
