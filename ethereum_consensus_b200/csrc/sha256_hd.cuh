// Portable (host + device) SHA-256 for hash_to_field's expand_message_xmd: a dozen compressions per message,
// so clarity beats speed here (the Merkle kernels use the tuned device version in sha256.cuh).
#pragma once
#include <cstddef>
#include <cstdint>

#include "fp.cuh"

namespace b200 {

struct Sha256Ctx {
    uint32_t st[8];
    uint8_t buf[64];
    uint32_t fill;
    uint64_t total;
};

B200_HD uint32_t sha_k(int i) {
    const uint32_t k[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
        0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
        0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
        0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
        0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
        0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
        0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return k[i];
}
B200_HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

B200_BIG void sha_compress(uint32_t st[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t(blk[4 * i]) << 24) | (uint32_t(blk[4 * i + 1]) << 16) | (uint32_t(blk[4 * i + 2]) << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll 1
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = h + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + sha_k(i) + w[i];
        uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
B200_HD void sha_init(Sha256Ctx& c) {
    c.st[0] = 0x6a09e667u; c.st[1] = 0xbb67ae85u; c.st[2] = 0x3c6ef372u; c.st[3] = 0xa54ff53au;
    c.st[4] = 0x510e527fu; c.st[5] = 0x9b05688cu; c.st[6] = 0x1f83d9abu; c.st[7] = 0x5be0cd19u;
    c.fill = 0; c.total = 0;
}
B200_HD void sha_update(Sha256Ctx& c, const uint8_t* d, size_t n) {
    for (size_t i = 0; i < n; i++) {
        c.buf[c.fill++] = d[i];
        if (c.fill == 64) { sha_compress(c.st, c.buf); c.fill = 0; }
    }
    c.total += n;
}
B200_HD void sha_final(Sha256Ctx& c, uint8_t out[32]) {
    uint64_t bits = c.total * 8;
    uint8_t pad = 0x80;
    sha_update(c, &pad, 1);
    pad = 0;
    while (c.fill != 56) sha_update(c, &pad, 1);
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[i] = uint8_t(bits >> (56 - 8 * i));
    sha_update(c, len, 8);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = uint8_t(c.st[i] >> 24); out[4 * i + 1] = uint8_t(c.st[i] >> 16);
        out[4 * i + 2] = uint8_t(c.st[i] >> 8); out[4 * i + 3] = uint8_t(c.st[i]);
    }
}

}  // namespace b200
