// CPU check of the pooled-storage Huffman dictionary build (imageencoder_b200/csrc/huffman.cu: build_dictionary) against the
// plain-allocator transcription of Huffman.cpp:237-311 it replaced (below, verbatim from the round-1 v9 build): same codes,
// same dictionary header bits, same return code, on real and random histograms (uniform, geometric, tie-heavy, few symbols).
// Test infrastructure, not product code.  Compiled as host code and linked against the product library only for the
// symbols huffman.cu's other functions reference; no device is touched.
#include "../../imageencoder_b200/csrc/huffman.cu"

#include <chrono>
#include <random>

namespace refimpl {
using namespace ie;
struct RefNode { uint8_t data; size_t freq; RefNode *left, *right; };
struct RefNodeCmp { bool operator()(const RefNode *a, const RefNode *b) const { return a->freq > b->freq; } };   // Huffman.hpp:70-74
struct RefCode { uint32_t word, len; };

void ref_walk(const RefNode *n, uint32_t word, uint32_t len, std::unordered_map<uint8_t, RefCode> &dict) {       // Huffman.cpp:79-104
    if (!n) return;
    if (!n->left && !n->right) { dict[n->data] = RefCode{word, len}; return; }
    ref_walk(n->left, word << 1, len + 1, dict);
    ref_walk(n->right, (word << 1) | 1u, len + 1, dict);
}
void ref_destroy(RefNode *n) { if (!n) return; ref_destroy(n->left); ref_destroy(n->right); delete n; }

struct RefBitWriter {
    std::vector<uint8_t> buf;
    size_t pos = 0;
    void put(unsigned len, uint32_t v) {
        for (unsigned p = 0; p < len; p++) {
            if ((pos >> 3) >= buf.size()) buf.resize(buf.size() + 256, 0);
            if ((v >> (len - 1 - p)) & 1u) buf[pos >> 3] |= uint8_t(1u << (7 - (pos & 7)));
            pos++;
        }
    }
};

// hist/first -> codes + dictionary header bits.  Returns IE_EINVAL when a code would exceed 32 bits (the reference's
// uint32 code words overflow there, Huffman.cpp:86-88).
static int ref_build_dictionary(const unsigned *hist, const unsigned long long *first, HuffCodes &codes, RefBitWriter &hdr) {
    std::vector<int> syms;
    for (int i = 0; i < 256; i++) if (hist[i]) syms.push_back(i);
    if (syms.empty()) { return IE_EINVAL; }
    std::sort(syms.begin(), syms.end(), [&](int a, int b) { return first[a] < first[b]; });
    std::unordered_map<uint8_t, uint32_t> freqs;                           // Huffman.cpp:237-243 (insertion = first occurrence)
    for (int s : syms) freqs[(uint8_t)s] = hist[s];
    std::priority_queue<RefNode *, std::vector<RefNode *>, RefNodeCmp> pq;       // Huffman.cpp:246-251
    for (const auto &pr : freqs) pq.push(new RefNode{pr.first, pr.second, nullptr, nullptr});
    while (pq.size() > 1) {                                                // Huffman.cpp:253-260
        RefNode *l = pq.top(); pq.pop();
        RefNode *r = pq.top(); pq.pop();
        pq.push(new RefNode{0xFF, l->freq + r->freq, l, r});
    }
    RefNode *root = pq.top();
    std::unordered_map<uint8_t, RefCode> dict;
    ref_walk(root, 0, 0, dict);                                                // Huffman.cpp:266
    ref_destroy(root);
    std::vector<std::pair<uint8_t, RefCode>> sorted(dict.begin(), dict.end());   // Huffman.cpp:269
    std::sort(sorted.begin(), sorted.end(),                                 // Huffman.cpp:272 (unstable, len descending)
              [](const std::pair<uint8_t, RefCode> &a, const std::pair<uint8_t, RefCode> &b) { return a.second.len > b.second.len; });
    std::unordered_map<uint32_t, uint32_t> bit_freqs;
    for (const auto &w : sorted) bit_freqs[w.second.len]++;
    uint32_t seq_len = 0, bit_len = 0;
    for (const auto &w : sorted) {                                         // Huffman.cpp:298-309
        if (w.second.len > 32) { return IE_EINVAL; }
        if (seq_len == 0) {
            bit_len = w.second.len;
            seq_len = bit_freqs[bit_len];
            hdr.put(8, 0x80u | (seq_len & 0x7Fu));                         // Huffman.cpp:36-46
            hdr.put(4, bit_len & 0xFu);
        }
        hdr.put(8, w.first);
        hdr.put(bit_len, w.second.word);
        seq_len--;
    }
    hdr.put(1, 0);                                                         // Huffman.cpp:311
    memset(&codes, 0, sizeof codes);
    for (const auto &pr : dict) { codes.word[pr.first] = pr.second.word; codes.len[pr.first] = (unsigned char)pr.second.len; }
    return IE_OK;
}


}  // namespace refimpl

int main(int argc, char **argv) {
    const long long n = (argc > 1) ? atoll(argv[1]) : 20000;
    std::mt19937_64 rng(99);
    long long fail = 0, too_long = 0;
    double t_new = 0, t_old = 0;
    for (long long it = 0; it < n; it++) {
        unsigned hist[256];
        unsigned long long first[256];
        const int mode = (int)(rng() % 6);
        const int nsym = (mode == 5) ? 1 + (int)(rng() % 4) : 1 + (int)(rng() % 256);
        int perm[256];
        for (int i = 0; i < 256; i++) perm[i] = i;
        for (int i = 255; i > 0; i--) std::swap(perm[i], perm[rng() % (i + 1)]);
        for (int i = 0; i < 256; i++) { hist[i] = 0; first[i] = ~0ull; }
        for (int k = 0; k < nsym; k++) {
            const int s = perm[k];
            unsigned h;
            switch (mode) {
            case 0: h = 1 + (unsigned)(rng() % 100000); break;                       // uniform counts
            case 1: h = 1 + (unsigned)(rng() % 4); break;                             // tie-heavy
            case 2: h = 1u << (rng() % 20); break;                                    // powers of two (ties between sums)
            case 3: h = 1 + (unsigned)(1e6 * std::exp(-0.05 * k)); break;            // geometric (long codes)
            case 4: h = 7; break;                                                     // all equal
            default: h = 1 + (unsigned)(rng() % 1000); break;
            }
            hist[s] = h;
            first[s] = (unsigned long long)k * 3 + (rng() % 3);                       // distinct first occurrences, order = perm order
        }
        HuffCodes ca, cb;
        HostBitWriter ha;
        refimpl::RefBitWriter hb;
        auto t0 = std::chrono::steady_clock::now();
        const int ra = ie::build_dictionary(hist, first, ca, ha);
        auto t1 = std::chrono::steady_clock::now();
        const int rb = refimpl::ref_build_dictionary(hist, first, cb, hb);
        auto t2 = std::chrono::steady_clock::now();
        t_new += std::chrono::duration<double, std::micro>(t1 - t0).count();
        t_old += std::chrono::duration<double, std::micro>(t2 - t1).count();
        bool ok = (ra == rb);
        if (ra == 0 && rb == 0) {
            ok = ok && ha.pos == hb.pos && memcmp(&ca, &cb, sizeof ca) == 0;
            for (size_t i = 0; ok && i < (ha.pos + 7) / 8; i++) ok = ha.buf[i] == hb.buf[i];
        } else {
            too_long++;
        }
        if (!ok) { if (fail < 10) printf("FAIL case %lld (mode %d, %d symbols): rc %d/%d, header bits %zu/%zu\n", it, mode, nsym, ra, rb, ha.pos, hb.pos); fail++; }
    }
    if (fail) { printf("huffdict_check: %lld failures\n", fail); return 1; }
    printf("huffdict_check: ok (%lld histograms, %lld with codes longer than 32 bits rejected by both; %.1f us per call, "
           "plain-allocator transcription %.1f us)\n", n, too_long, t_new / n, t_old / n);
    return 0;
}
