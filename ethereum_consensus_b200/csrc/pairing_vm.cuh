// Lane-parallel "Fp2 VM": executes the statically scheduled Miller-loop / final-exponentiation programs produced by
// tools/gen_pairing_vm.py.  A TEAM of kVmTeam lanes shares one register file of Fp2 slots (shared memory on the
// device); in every round each lane executes at most one Fp2 instruction, rounds are separated by a barrier.
// Register allocation guarantees that no slot is both read and written in the same round.
#pragma once
#include "fp12.cuh"
#include "pairing_vm_prog.cuh"  // teams of 8 lanes (teams of 16 measured slower on B200: profiles/r1_tuning.md)

namespace b200 {

// opcodes (must match tools/gen_pairing_vm.py).  VM_LIN: r = sum of up to four terms (+-1, +-2, +-3) * (1 or xi) * slot —
// a whole single-use chain of additions / subtractions / doublings / xi-multiplications in ONE instruction (round 2: the
// generator fuses such chains; light ops then cost one round per chain, not per op).  VM_FMUL / VM_FSQR: products whose
// operands are (s0 +- s1) and (s2 +- s3) — the Karatsuba pre-additions ride in the product's own round.
enum VmOp : uint32_t { VM_NOP = 0, VM_MUL, VM_SQR, VM_MULFP, VM_INV, VM_ADD, VM_SUB, VM_NEG, VM_DBL, VM_CONJ, VM_MULXI, VM_COPY, VM_LDC,
                       VM_LIN, VM_FMUL, VM_FSQR };

// register-file accessors: dense array of Fp2 (host) or 25-word-strided shared memory (device, bank-conflict free)
struct VmRfDense {
    Fp2* p;
    B200_HD Fp2 load(uint32_t i) const { return p[i]; }
    B200_HD void store(uint32_t i, const Fp2& v) const { p[i] = v; }
};
// Slot stride in words.  25 (round 1): scalar LDS/STS, conflict-free for any slot pattern.  24 / 28 (-DB200_VM_SLOT_WORDS=..):
// 16-byte aligned slots moved with 128-bit LDS/STS — 6 + 6 + 6 wide accesses per light op instead of 24 + 24 + 24.
#if !defined(B200_VM_SLOT_WORDS)
#define B200_VM_SLOT_WORDS 25
#endif
constexpr int kVmSlotWords = B200_VM_SLOT_WORDS;
struct VmRfStrided {
    uint32_t* p;
    B200_HD Fp2 load(uint32_t i) const {
        Fp2 v;
#if defined(__CUDA_ARCH__) && (B200_VM_SLOT_WORDS % 4 == 0)
        const uint4* q = reinterpret_cast<const uint4*>(p + i * kVmSlotWords);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const uint4 a = q[k], b = q[3 + k];
            v.c0.l[4 * k] = a.x; v.c0.l[4 * k + 1] = a.y; v.c0.l[4 * k + 2] = a.z; v.c0.l[4 * k + 3] = a.w;
            v.c1.l[4 * k] = b.x; v.c1.l[4 * k + 1] = b.y; v.c1.l[4 * k + 2] = b.z; v.c1.l[4 * k + 3] = b.w;
        }
#else
        const uint32_t* q = p + i * kVmSlotWords;
#pragma unroll
        for (int k = 0; k < 12; k++) { v.c0.l[k] = q[k]; v.c1.l[k] = q[12 + k]; }
#endif
        return v;
    }
    B200_HD void store(uint32_t i, const Fp2& v) const {
#if defined(__CUDA_ARCH__) && (B200_VM_SLOT_WORDS % 4 == 0)
        uint4* q = reinterpret_cast<uint4*>(p + i * kVmSlotWords);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            q[k] = make_uint4(v.c0.l[4 * k], v.c0.l[4 * k + 1], v.c0.l[4 * k + 2], v.c0.l[4 * k + 3]);
            q[3 + k] = make_uint4(v.c1.l[4 * k], v.c1.l[4 * k + 1], v.c1.l[4 * k + 2], v.c1.l[4 * k + 3]);
        }
#else
        uint32_t* q = p + i * kVmSlotWords;
#pragma unroll
        for (int k = 0; k < 12; k++) { q[k] = v.c0.l[k]; q[12 + k] = v.c1.l[k]; }
#endif
    }
};

// one LIN / pre-add term: code = sign 8 | xi 4 | magnitude 1..3 (0 = absent)
template <class RF>
B200_HD void vm_term(Fp2& acc, bool& have, const RF& rf, uint32_t slot, uint32_t code) {
    if (!(code & 3u)) return;
    Fp2 x = rf.load(slot);
    if (code & 4u) { Fp2 t; fp2_mul_xi(t, x); x = t; }
    Fp2 t = x;
    if ((code & 3u) >= 2u) fp2_dbl(t, x);
    if ((code & 3u) == 3u) fp2_add(t, t, x);
    if (!have) { if (code & 8u) fp2_neg(acc, t); else acc = t; have = true; }
    else if (code & 8u) fp2_sub(acc, acc, t);
    else fp2_add(acc, acc, t);
}

// one instruction (two words, see pairing_vm_prog.cuh): result into `res`, returns false for NOP
template <class RF>
B200_HD bool vm_exec(uint32_t w0, uint32_t w1, const RF& rf, const Fp2* consts, Fp2& res) {
    const uint32_t op = w0 & 0xffu, s0 = (w0 >> 16) & 0xffu, s1 = w0 >> 24, s2 = w1 & 0xffu, s3 = (w1 >> 8) & 0xffu;
    const uint32_t c0 = (w1 >> 16) & 0xfu, c1 = (w1 >> 20) & 0xfu, c2 = (w1 >> 24) & 0xfu, c3 = w1 >> 28;
    switch (op) {
    case VM_NOP: return false;
    case VM_MUL: { const Fp2 x = rf.load(s0), y = rf.load(s1); fp2_mul(res, x, y); } break;
    case VM_SQR: { const Fp2 x = rf.load(s0); fp2_sqr(res, x); } break;
    case VM_MULFP: { const Fp2 x = rf.load(s0); const Fp k = rf.load(s1).c0; fp2_mul_fp(res, x, k); } break;
    case VM_INV: { const Fp2 x = rf.load(s0); fp2_inv(res, x); } break;
    case VM_LIN: {
        bool have = false;
        vm_term(res, have, rf, s0, c0); vm_term(res, have, rf, s1, c1); vm_term(res, have, rf, s2, c2); vm_term(res, have, rf, s3, c3);
    } break;
    case VM_FMUL: {
        Fp2 l, r;
        bool hl = false, hr = false;
        vm_term(l, hl, rf, s0, c0); vm_term(l, hl, rf, s1, c1);
        vm_term(r, hr, rf, s2, c2); vm_term(r, hr, rf, s3, c3);
        fp2_mul(res, l, r);
    } break;
    case VM_FSQR: {
        Fp2 l;
        bool hl = false;
        vm_term(l, hl, rf, s0, c0); vm_term(l, hl, rf, s1, c1);
        fp2_sqr(res, l);
    } break;
    case VM_CONJ: { const Fp2 x = rf.load(s0); fp2_conj(res, x); } break;
    case VM_COPY: res = rf.load(s0); break;
    case VM_LDC: res = consts[s0]; break;
    default: return false;   // the unfused light opcodes are no longer emitted
    }
    return true;
}

// Sequential reference executor (host tests): all lanes of a round read before any writes.
inline void vm_run_host(const uint32_t* code, int n_rounds, const Fp2* consts, Fp2* rf_mem) {
    VmRfDense rf{rf_mem};
    for (int r = 0; r < n_rounds; r++) {
        Fp2 res[kVmTeam];
        bool live[kVmTeam];
        for (int l = 0; l < kVmTeam; l++) live[l] = vm_exec(code[2 * (r * kVmTeam + l)], code[2 * (r * kVmTeam + l) + 1], rf, consts, res[l]);
        for (int l = 0; l < kVmTeam; l++)
            if (live[l]) rf.store((code[2 * (r * kVmTeam + l)] >> 8) & 0xffu, res[l]);
    }
}

// plain-integer constant table -> Montgomery form
B200_HD void vm_const_to_mont(Fp2& out, const uint32_t limbs[24]) {
    Fp a, b;
#pragma unroll
    for (int i = 0; i < 12; i++) { a.l[i] = limbs[i]; b.l[i] = limbs[12 + i]; }
    fp_to_mont(out.c0, a);
    fp_to_mont(out.c1, b);
}

}  // namespace b200
