// TEST INFRASTRUCTURE (CPU): equivalence of the two "push the surviving internal children" variants of
// gfxexp_b200/csrc/traverse.cuh (the loop, and the branch-free GFX_TRAVERSE_PREDICATED_PUSH one), transcribed with
// __popc/__ffs -> builtins, on random sorted key sets incl. ties, culled children and a nearly full stack.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>
#include <algorithm>
#include <vector>
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
constexpr int kStackSize = 64;
struct St { uint32_t nodeIdx; int sp; uint32_t stack[kStackSize][2]; float best; bool overflow; };
static bool loopVariant(const uint32_t keys[8], uint32_t childBase, uint32_t internalMask, St &st) {
    uint32_t next = 0xFFFFFFFFu, nextT = 0u;
    for (int k = 7; k >= 0; --k) {
        const uint32_t key = keys[k];
        if (key == 0xFFFFFFFFu || key < 0x80000000u) continue;
        const uint32_t slot = key & 7u;
        const uint32_t tnBits = (key & 0x7FFFFFF8u) << 1;
        if (u2f(tnBits) > st.best) continue;
        if (next != 0xFFFFFFFFu) {
            if (st.sp < kStackSize) { st.stack[st.sp][0] = next; st.stack[st.sp][1] = nextT; st.sp++; }
            else st.overflow = true;
        }
        next = childBase + __builtin_popcount(internalMask & ((1u << slot) - 1u));
        nextT = tnBits;
    }
    if (next != 0xFFFFFFFFu) { st.nodeIdx = next; return true; }
    return false;
}
static bool predVariant(const uint32_t keys[8], uint32_t childBase, uint32_t internalMask, St &st) {
    uint32_t survivors = 0u, childNode[8], childT[8];
    for (int k = 0; k < 8; ++k) {
        const uint32_t key = keys[k];
        const uint32_t slot = key & 7u;
        childT[k] = (key & 0x7FFFFFF8u) << 1;
        childNode[k] = childBase + __builtin_popcount(internalMask & ((1u << slot) - 1u));
        const bool survives = key != 0xFFFFFFFFu && key >= 0x80000000u && !(u2f(childT[k]) > st.best);
        survivors |= survives ? (1u << k) : 0u;
    }
    if (survivors) {
        const int nearest = __builtin_ffs(survivors) - 1;
        for (int k = 7; k >= 1; --k) {
            const uint32_t at = (uint32_t)st.sp + __builtin_popcount(survivors >> (k + 1));
            if (((survivors >> k) & 1u) && k != nearest) {
                if (at < (uint32_t)kStackSize) { st.stack[at][0] = childNode[k]; st.stack[at][1] = childT[k]; }
                else st.overflow = true;
            }
        }
        st.sp = std::min(st.sp + __builtin_popcount(survivors) - 1, kStackSize);
        uint32_t next = childNode[0];
        for (int k = 1; k < 8; ++k) next = k == nearest ? childNode[k] : next;
        st.nodeIdx = next;
        return true;
    }
    return false;
}
int main() {
    std::mt19937 rng(12345);
    std::uniform_real_distribution<float> dist(0.0f, 50.0f);
    long mismatches = 0, cases = 0;
    for (int it = 0; it < 4000000; ++it) {
        // random node: each slot is a miss, a leaf hit or an internal hit with an entry distance; keys sorted ascending
        uint32_t keys[8]; uint32_t internalMask = rng() & 0xFFu;
        for (int slot = 0; slot < 8; ++slot) {
            const uint32_t r = rng() % 10;
            if (r < 3) { keys[slot] = 0xFFFFFFFFu; continue; }
            const float tn = (rng() % 7 == 0) ? 10.0f : dist(rng); // ties on purpose
            const uint32_t isInternal = (internalMask >> slot) & 1u;
            keys[slot] = (isInternal << 31) | ((f2u(tn) >> 1) & 0x7FFFFFF8u) | (uint32_t)slot;
        }
        std::sort(keys, keys + 8);
        St a{}, b{};
        a.sp = b.sp = (it % 50 == 0) ? kStackSize - (int)(rng() % 4) : (int)(rng() % 40);
        a.best = b.best = (rng() % 4 == 0) ? 3.4e38f : dist(rng);
        for (int i = 0; i < kStackSize; ++i) { a.stack[i][0] = b.stack[i][0] = 7777u + i; a.stack[i][1] = b.stack[i][1] = 0; }
        const uint32_t childBase = rng() % 100000;
        const bool ra = loopVariant(keys, childBase, internalMask, a), rb = predVariant(keys, childBase, internalMask, b);
        ++cases;
        bool same = ra == rb && a.sp == b.sp && a.overflow == b.overflow && (!ra || a.nodeIdx == b.nodeIdx);
        for (int i = 0; same && i < std::min(a.sp, kStackSize); ++i)
            same = a.stack[i][0] == b.stack[i][0] && a.stack[i][1] == b.stack[i][1];
        if (!same && ++mismatches < 5)
            printf("mismatch: ra %d rb %d sp %d/%d node %u/%u ovf %d/%d\n", ra, rb, a.sp, b.sp, a.nodeIdx, b.nodeIdx, a.overflow, b.overflow);
    }
    printf("%ld cases, %ld mismatches\n", cases, mismatches);
    return mismatches != 0;
}
