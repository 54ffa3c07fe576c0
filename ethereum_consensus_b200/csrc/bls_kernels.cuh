// Device buffers and kernel launchers of the BLS batch-verification pipeline (see bls_engine.cu for the flow).
#pragma once
#include <cstddef>
#include <cstdint>

#include "groups.cuh"
#include "fp12.cuh"

namespace b200 {

// G1 operand of the lane-parallel Miller loop: (X Z, Y, Z^3) of a Jacobian point (affine point: (x, y, 1)); the line
// functions absorb the Z^3 scaling, so the aggregate key is never inverted
struct G1Pre {
    Fp xz, y, z3;
    uint32_t inf;
};

// per-tuple flags written by the G1 aggregation kernel
enum : uint32_t { TUPLE_FLAG_EMPTY = 1u, TUPLE_FLAG_AGG_INF = 2u };
// per-signature status written by the signature kernel
enum : int32_t { SIG_OK = 0, SIG_NOT_IN_GROUP = -1 };  // >0: blst decode error code

// K1: key_validate every 48-byte public key -> affine point + blst code
void set_g1_variant(int v);
void set_g1_small_n(uint32_t n);
void set_small_cta(int threads);
void set_vm_team16_max(uint32_t n);
void set_vm_cta(int threads);
// cta = 0: by batch size (384-thread CTAs above the small-batch bound, 128 below); 128 / 384 force one (default variant only)
void launch_g1_validate(const uint8_t* keys, uint32_t n, G1Aff* out, int32_t* codes, void* stream, int cta = 0);
// K2: per tuple t, sum the validated keys [off[t], off[t+1]) (or gather through `index` when non-null);
//     first failing key (in order) decides pk_code[t]
//     agg == nullptr: only the code scan (aggregate_verify keeps keys separate); extra_flags is OR-ed into flags
//     agg_pre != nullptr: write the un-normalised sum for the VM Miller kernel instead of the affine point
void launch_g1_aggregate(const G1Aff* keys, const int32_t* key_codes, const uint32_t* index, const uint32_t* off,
                         uint32_t n_tuples, G1Aff* agg, G1Pre* agg_pre, int32_t* pk_code, uint32_t* flags,
                         uint32_t extra_flags, void* stream, G1Jac* agg_jac = nullptr);
// RLC whole-batch check (bls_rlc.cu): scale per tuple, fold 32 -> 1 per launch, sum -> affine
void launch_rlc_scale(const G1Jac* agg, const G2Aff* sig, const int32_t* pk_code, const uint32_t* flags, const int32_t* sig_code,
                      const uint32_t* seed_words, uint64_t t0, uint32_t n, G1Pre* out_g1, G2Jac* out_g2, int32_t* bad, void* stream);
uint32_t launch_rlc_reduce(const Fp12* f_in, const G2Jac* q_in, uint32_t n, Fp12* f_out, G2Jac* q_out, void* stream);
void launch_rlc_finish(const G2Jac* q, G2Aff* out, void* stream);
void launch_fp12_one(Fp12* out, void* stream);
// K3: decompress + subgroup-check every 96-byte signature
//     `threads`: CTA size (32 = spread for latency, 512 = pack onto few SMs while the per-key kernel runs)
void launch_g2_sig_decode(const uint8_t* sigs, uint32_t n, G2Aff* out, int32_t* sig_code, void* stream);
// K4: hash_to_G2 of message i = bytes [moff[i], moff[i+1]) of `msgs`
//     `tmp_jac`: scratch for 2n Jacobian G2 points (288 B each)
void launch_hash_to_g2(const uint8_t* msgs, const uint32_t* moff, uint32_t n, G2Aff* out, void* tmp_jac, void* stream);
// K5: one Miller loop per pair (g1[g1_idx[i]], g2[g2_idx[i]]); pairs whose tuple already failed are skipped
void launch_miller(const G1Aff* g1, const uint32_t* g1_idx, const G2Aff* g2, const uint32_t* g2_idx,
                   const uint32_t* pair_tuple, const int32_t* pk_code, const uint32_t* flags, const int32_t* sig_code,
                   uint32_t n_pairs, Fp12* f, void* stream);
// K6: per tuple: merge codes with the reference's precedence, multiply its Miller values, final exponentiation
void launch_final(const Fp12* f, const uint32_t* pair_off, const int32_t* pk_code, const uint32_t* flags,
                  const int32_t* sig_code, uint32_t n_tuples, int32_t* out_codes, void* stream);
// lane-parallel (team) versions of K5 / K6 for tuples with exactly two pairs (bls_vm.cu); vm_init returns 0 on success
int vm_init(void* stream);
// replaces one team size's scheduled programs (blob layout: bls_vm.cu); returns 0 on success, 1 on a malformed blob
int vm_load_programs(const uint32_t* blob, size_t n_words, void* stream);
void launch_vm_miller(const G1Pre* g1, const uint32_t* g1_idx, const G2Aff* g2, const uint32_t* g2_idx,
                      const uint32_t* pair_tuple, const int32_t* pk_code, const uint32_t* flags, const int32_t* sig_code,
                      uint32_t n_pairs, Fp12* f, void* stream);
void launch_vm_final(const Fp12* f, const uint32_t* pair_off, const int32_t* pk_code, const uint32_t* flags,
                     const int32_t* sig_code, uint32_t n_tuples, int32_t* out_codes, void* stream);
// aggregation helpers for `aggregate` / `eth_aggregate_public_keys`
void launch_g2_sum_compress(const G2Aff* sigs, const int32_t* sig_code, uint32_t n, uint8_t* out96, int32_t* out_code, void* stream);
void launch_g1_compress(const G1Aff* p, uint8_t* out48, void* stream);
// writes -g1 (the negated generator) to *out (and its G1Pre form)
void launch_neg_g1(G1Aff* out, G1Pre* out_pre, void* stream);
// on-device self-test of Fp arithmetic (portable vs tuned paths), returns mismatches in *out
void launch_fp_selftest(uint32_t n, uint32_t seed, uint32_t* out_mismatch, void* stream);

}  // namespace b200
