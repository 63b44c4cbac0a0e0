// Emulation of libc++'s std::unordered_map<uint32_t, uint64_t> iteration order (identity hash), from memory of __hash_table.
#pragma once
#include <cstdint>
#include <cmath>
#include <vector>
#include <utility>
#include <stdexcept>
struct LibcxxMap {
    struct Node { std::pair<uint32_t, uint64_t> kv; int next; };
    std::vector<Node> nodes;      // index 0 = before-begin sentinel
    std::vector<int> buckets;     // predecessor node index of the bucket's first node, -1 = empty
    size_t sz = 0;
    LibcxxMap() { nodes.push_back({{0, 0}, -1}); }
    static bool is_pow2(size_t bc) { return bc > 2 && !(bc & (bc - 1)); }
    static size_t constrain(size_t h, size_t bc) { return !(bc & (bc - 1)) ? h & (bc - 1) : (h < bc ? h : h % bc); }
    static bool is_prime(size_t n) { if (n < 2) return false; for (size_t d = 2; d * d <= n; d++) if (n % d == 0) return false; return true; }
    static size_t next_prime(size_t n) { while (!is_prime(n)) n++; return n; }
    void do_rehash(size_t nbc) {
        buckets.assign(nbc, -1);
        int pp = 0, cp = nodes[0].next;
        if (cp < 0) return;
        size_t chash = constrain(nodes[cp].kv.first, nbc);
        buckets[chash] = pp;
        size_t phash = chash;
        for (pp = cp, cp = nodes[cp].next; cp >= 0; cp = nodes[pp].next) {
            chash = constrain(nodes[cp].kv.first, nbc);
            if (chash == phash) pp = cp;
            else if (buckets[chash] < 0) { buckets[chash] = pp; pp = cp; phash = chash; }
            else {
                int np = cp;
                nodes[pp].next = nodes[np].next;
                nodes[np].next = nodes[buckets[chash]].next;
                nodes[buckets[chash]].next = cp;
            }
        }
    }
    void rehash(size_t n) {
        if (n == 1) n = 2; else if (n & (n - 1)) n = next_prime(n);
        if (n > buckets.size()) do_rehash(n);
    }
    int find_idx(uint32_t k) const {
        size_t bc = buckets.size();
        if (!bc) return -1;
        size_t chash = constrain(k, bc);
        int nd = buckets[chash];
        if (nd < 0) return -1;
        for (nd = nodes[nd].next; nd >= 0 && constrain(nodes[nd].kv.first, bc) == chash; nd = nodes[nd].next)
            if (nodes[nd].kv.first == k) return nd;
        return -1;
    }
    uint64_t &operator[](uint32_t k) {
        int f = find_idx(k);
        if (f >= 0) return nodes[f].kv.second;
        size_t bc = buckets.size();
        if (sz + 1 > bc || bc == 0) {
            rehash(std::max<size_t>(2 * bc + !is_pow2(bc), (size_t)std::ceil(float(sz + 1) / 1.0f)));
            bc = buckets.size();
        }
        size_t chash = constrain(k, bc);
        nodes.push_back({{k, 0}, -1});
        int h = (int)nodes.size() - 1;
        int pn = buckets[chash];
        if (pn < 0) {
            pn = 0;
            nodes[h].next = nodes[pn].next;
            nodes[pn].next = h;
            buckets[chash] = pn;
            if (nodes[h].next >= 0) buckets[constrain(nodes[nodes[h].next].kv.first, bc)] = h;
        } else {
            nodes[h].next = nodes[pn].next;
            nodes[pn].next = h;
        }
        sz++;
        return nodes[h].kv.second;
    }
    void emplace(uint32_t k, uint64_t v) { if (find_idx(k) < 0) (*this)[k] = v; }
    void reserve(size_t) {}
    struct It {
        const LibcxxMap *m; int i;
        bool operator!=(const It &o) const { return i != o.i; }
        bool operator==(const It &o) const { return i == o.i; }
        void operator++() { i = m->nodes[i].next; }
        const std::pair<uint32_t, uint64_t> &operator*() const { return m->nodes[i].kv; }
        const std::pair<uint32_t, uint64_t> *operator->() const { return &m->nodes[i].kv; }
    };
    It begin() const { return {this, nodes[0].next}; }
    It end() const { return {this, -1}; }
    It find(uint32_t k) const { return {this, find_idx(k)}; }
    uint64_t at(uint32_t k) const { int f = find_idx(k); if (f < 0) throw std::out_of_range("at"); return nodes[f].kv.second; }
    size_t size() const { return sz; }
};
