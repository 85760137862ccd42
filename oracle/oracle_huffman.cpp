/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the byte-wise Huffman stage (Huffman.cpp / Huffman.hpp of the reference).
 *
 * This one file is C++ rather than C on purpose: the reference's output depends on the iteration order of
 * std::unordered_map<uint8_t,...>, on the tie behaviour of std::priority_queue and on std::sort (unstable) --
 * i.e. on libstdc++ itself.  The oracle therefore states the algorithm with the same
 * libstdc++ containers, fed in the same order, instead of re-modelling libstdc++ internals.  Only the container
 * sequence is shared with the reference; bit I/O, tree and dictionary handling are restated from the description
 * in Huffman.cpp:232-344 (encode) and :119-204, :354-402 (decode).
 *
 * Parity status: PINNED against oracle/_ref (the compiled reference) by tests/test_oracle_vs_reference.py.
 * Nothing in the product may include, link or call this file.
 */
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <queue>
#include <unordered_map>
#include <vector>

namespace {

struct HNode {
    int sym;            // -1 for internal nodes (the reference stores uint8_t(-1) = 255; never inspected for internals)
    size_t freq;
    HNode *l, *r;
};
struct HNodeGreater {   // Huffman.hpp:70-74
    bool operator()(const HNode *a, const HNode *b) const { return a->freq > b->freq; }
};
struct Code { uint32_t word, len; };

struct BitW {
    std::vector<uint8_t> buf;
    size_t pos = 0;
    void put(unsigned len, uint32_t v) {
        for (unsigned p = 0; p < len; p++) {
            if ((pos >> 3) >= buf.size()) buf.resize(buf.size() + buf.size() / 2 + 64, 0);
            if ((v >> (len - 1 - p)) & 1u) buf[pos >> 3] |= uint8_t(1u << (7 - (pos & 7)));
            pos++;
        }
    }
};

void build_dict(const HNode *n, uint32_t word, uint32_t len, std::unordered_map<uint8_t, Code> &dict) {   // Huffman.cpp:79-104
    if (!n) return;
    if (!n->l && !n->r) { dict[uint8_t(n->sym)] = Code{word, len}; return; }
    build_dict(n->l, (word << 1), len + 1, dict);          // left  = '0'
    build_dict(n->r, (word << 1) | 1u, len + 1, dict);     // right = '1'
}
void free_tree(HNode *n) { if (!n) return; free_tree(n->l); free_tree(n->r); delete n; }

}  // namespace

extern "C" {

/* Huffman.cpp:232-344.  in = the byte-rounded plain stream.  Returns the number of output BYTES
 * (round_to_byte of the bit position, as saveResult writes it), or -1 if out_cap is too small.
 * code_len_out / code_word_out (optional, 256 entries each): the dictionary, len 0xFFFFFFFF = symbol absent.
 * reverted_out (optional): 1 if the "no extra compression" branch was taken (Huffman.cpp:329-341). */
long long orc_huffman_encode(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap,
                             uint32_t *code_len_out, uint32_t *code_word_out, int *reverted_out) {
    std::unordered_map<uint8_t, uint32_t> freqs;                       // Huffman.cpp:237-243
    for (size_t i = 0; i < n; i++) freqs[in[i]]++;

    std::priority_queue<HNode *, std::vector<HNode *>, HNodeGreater> pq;   // Huffman.cpp:246-251
    for (const auto &p : freqs) pq.push(new HNode{p.first, p.second, nullptr, nullptr});
    if (pq.empty()) return -3;
    while (pq.size() > 1) {                                            // Huffman.cpp:253-260
        HNode *l = pq.top(); pq.pop();
        HNode *r = pq.top(); pq.pop();
        pq.push(new HNode{-1, l->freq + r->freq, l, r});
    }
    HNode *root = pq.top();

    std::unordered_map<uint8_t, Code> dict;                            // Huffman.cpp:266
    build_dict(root, 0, 0, dict);
    std::vector<std::pair<uint8_t, Code>> sorted(dict.begin(), dict.end());   // Huffman.cpp:269
    std::sort(sorted.begin(), sorted.end(),                            // Huffman.cpp:272, comparator :15-21
              [](const std::pair<uint8_t, Code> &a, const std::pair<uint8_t, Code> &b) { return a.second.len > b.second.len; });
    std::unordered_map<uint32_t, uint32_t> bit_freqs;                  // Huffman.cpp:275-278
    for (const auto &w : sorted) bit_freqs[w.second.len]++;

    BitW w;
    w.buf.assign(n + 1024, 0);
    uint32_t seq_len = 0, bit_len = 0;
    for (const auto &e : sorted) {                                     // Huffman.cpp:298-309
        if (seq_len == 0) {
            bit_len = e.second.len;
            seq_len = bit_freqs[bit_len];
            if (seq_len > 0) {                                         // Huffman.cpp:36-46
                w.put(8, 0x80u | (seq_len & 0x7Fu));
                w.put(4, bit_len & 0xFu);
            } else {
                w.put(1, 0);
            }
        }
        w.put(8, e.first);
        w.put(bit_len, e.second.word);
        seq_len--;
    }
    w.put(1, 0);                                                       // stop bit, Huffman.cpp:311
    Code lut[256];
    for (int i = 0; i < 256; i++) lut[i] = Code{0, 0xFFFFFFFFu};
    for (const auto &p : dict) lut[p.first] = p.second;
    if (code_len_out) for (int i = 0; i < 256; i++) code_len_out[i] = lut[i].len;
    if (code_word_out) for (int i = 0; i < 256; i++) code_word_out[i] = lut[i].word;
    for (size_t i = 0; i < n; i++) w.put(lut[in[i]].len, lut[in[i]].word);   // Huffman.cpp:314-319
    free_tree(root);

    size_t total = (w.pos + 7) / 8;
    int reverted = 0;
    if (n < total) {                                                   // Huffman.cpp:329-341: '0' bit + the input bytes
        reverted = 1;
        BitW r;
        r.buf.assign(n + 2, 0);
        r.put(1, 0);
        for (size_t i = 0; i < n; i++) r.put(8, in[i]);
        w = r;
        total = (w.pos + 7) / 8;
    }
    if (reverted_out) *reverted_out = reverted;
    if (total > out_cap) return -1;
    std::memcpy(out, w.buf.data(), total);
    return (long long)total;
}

/* Huffman.cpp:354-402 + :119-180 + :190-204.  Returns the number of bytes in out; *start_bit_out = bit position at
 * which the plain stream starts inside out (1 for pass-through, 0 after real decoding).  -1: out_cap too small,
 * -2: malformed dictionary. */
long long orc_huffman_decode(const uint8_t *in, size_t n, uint8_t *out, size_t out_cap, size_t *start_bit_out) {
    size_t pos = 0;
    const size_t raw_bits = n * 8;
    auto get_bit = [&]() -> uint32_t {
        if ((pos >> 3) >= n) return 0;                                 // BitStream.cpp:17-20 (no advance past the end)
        uint32_t b = (in[pos >> 3] >> (7 - (pos & 7))) & 1u; pos++; return b;
    };
    auto get = [&](unsigned l) -> uint32_t { uint32_t v = 0; for (unsigned i = 0; i < l; i++) v |= get_bit() << (l - i - 1); return v; };

    struct DN { int sym; int l, r; };
    std::vector<DN> tree(1, DN{-1, -1, -1});
    bool any = false;
    while (get_bit()) {                                                // Huffman.cpp:57-65, 128-142
        uint32_t seq = get(7), bl = get(4);
        while (seq--) {
            uint8_t key = uint8_t(get(8));
            uint32_t word = get(bl);
            if (bl == 0) return -2;                                    // the reference shifts by -1 here (UB)
            int cur = 0;                                               // Huffman.cpp:153-180
            for (int b = int(bl) - 1; b >= 0; b--) {
                bool right = (word >> b) & 1u;
                int &child = right ? tree[cur].r : tree[cur].l;
                if (b == 0) { tree.push_back(DN{key, -1, -1}); (right ? tree[cur].r : tree[cur].l) = int(tree.size()) - 1; }
                else {
                    if (child < 0) { tree.push_back(DN{-1, -1, -1}); (right ? tree[cur].r : tree[cur].l) = int(tree.size()) - 1; }
                    cur = right ? tree[cur].r : tree[cur].l;
                }
            }
            any = true;
        }
    }
    if (!any) {                                                        // Huffman.cpp:361-371: pass-through
        if (n > out_cap) return -1;
        std::memcpy(out, in, n);
        *start_bit_out = pos;                                          // == 1
        return (long long)n;
    }
    size_t o = 0;
    while (pos < raw_bits) {                                           // Huffman.cpp:376-383
        int cur = 0;
        size_t guard = 0;
        while (tree[cur].l >= 0 || tree[cur].r >= 0) {
            int nxt = get_bit() ? tree[cur].r : tree[cur].l;
            if (nxt < 0 || ++guard > 64) return -2;                    // the reference dereferences nullptr here
            cur = nxt;
        }
        if (o >= out_cap) return -1;
        out[o++] = uint8_t(tree[cur].sym);
    }
    *start_bit_out = 0;
    return (long long)o;
}

}  // extern "C"
