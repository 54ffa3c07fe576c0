// Lane-parallel pairing kernels: a TEAM of 8 or 16 lanes per (G1, G2) pair / per tuple replays the statically
// scheduled Fp2 programs of tools/gen_pairing_vm.py (Miller loop, final exponentiation) on a shared-memory register
// file.  Replaces the one-thread-per-pair kernels of bls_pairing.cu on the batch path: those expose only 2T threads
// (a ~40 ms latency floor at T = 4096); here 16x more lanes work on the same tuples, products stay inlined PTX.
#include <cuda_runtime.h>

#include <cstdlib>

#include "bls_kernels.cuh"
#include "pairing_vm.cuh"

namespace b200 {
namespace {

__global__ void k_vm_consts(const uint32_t* __restrict__ plain, Fp2* __restrict__ out) {
    const int i = threadIdx.x;
    if (i < kVmConsts) {
        uint32_t l[24];
        for (int k = 0; k < 24; k++) l[k] = plain[i * 24 + k];
        Fp2 v;
        vm_const_to_mont(v, l);
        out[i] = v;
    }
}

template <int TEAM>
__device__ __forceinline__ void vm_run(const uint32_t* __restrict__ code, int n_rounds, const Fp2* __restrict__ consts,
                                       const VmRfStrided& rf, uint32_t lane, bool active) {
    // the next round's instruction word is fetched while the current round executes: with one or two warps per scheduler
    // (small batches) the load's latency would otherwise sit on the critical path of every one of the ~2 500 / ~3 800 rounds
    uint32_t w = __ldg(code + lane);
#pragma unroll 1
    for (int r = 0; r < n_rounds; r++) {
        const uint32_t w_next = (r + 1 < n_rounds) ? __ldg(code + (r + 1) * TEAM + lane) : 0u;
        if (active && (w & 0xffu) != VM_NOP) {
            Fp2 res;
            vm_exec(w, rf, consts, res);
            rf.store((w >> 8) & 0xffu, res);
        }
        __syncwarp();
        w = w_next;
    }
}

__device__ __forceinline__ bool tuple_dead(uint32_t t, const int32_t* pk_code, const uint32_t* flags, const int32_t* sig_code) {
    return pk_code[t] != BLS_SUCCESS || flags[t] != 0 || sig_code[t] != SIG_OK;
}

// one team per pair
template <int TEAM>
__global__ void __launch_bounds__(128) k_vm_miller(const uint32_t* __restrict__ code, const Fp2* __restrict__ consts,
                                                    const G1Pre* __restrict__ g1, const uint32_t* __restrict__ g1_idx,
                                                    const G2Aff* __restrict__ g2, const uint32_t* __restrict__ g2_idx,
                                                    const uint32_t* __restrict__ pair_tuple, const int32_t* __restrict__ pk_code,
                                                    const uint32_t* __restrict__ flags, const int32_t* __restrict__ sig_code,
                                                    uint32_t n_pairs, Fp12* __restrict__ f) {
    typedef VmProg<TEAM> P;
    extern __shared__ uint32_t smem[];
    const uint32_t team_in_block = threadIdx.x / TEAM, lane = threadIdx.x % TEAM;
    const uint32_t i = blockIdx.x * (blockDim.x / TEAM) + team_in_block;
    VmRfStrided rf{smem + team_in_block * (P::miller_slots * kVmSlotWords)};
    bool active = i < n_pairs && !tuple_dead(pair_tuple[i], pk_code, flags, sig_code);
    bool trivial = false;  // a point at infinity: the pair contributes 1
    if (active) {
        const G1Pre* p = g1 + g1_idx[i];
        const G2Aff* q = g2 + g2_idx[i];
        trivial = p->inf || q->inf;
        if (!trivial) {  // program inputs: slots 0..4 = X Z, Y, Z^3 (as Fp2 with c1 = 0), Qx, Qy
            if (lane == 0) { Fp2 v; v.c0 = p->xz; v.c1 = fp_zero(); rf.store(0, v); }
            if (lane == 1) { Fp2 v; v.c0 = p->y; v.c1 = fp_zero(); rf.store(1, v); }
            if (lane == 2) { Fp2 v; v.c0 = p->z3; v.c1 = fp_zero(); rf.store(2, v); }
            if (lane == 3) rf.store(3, q->x);
            if (lane == 4) rf.store(4, q->y);
        }
    }
    __syncwarp();
    vm_run<TEAM>(code, P::miller_rounds, consts, rf, lane, active && !trivial);
    static_assert(TEAM >= 6, "team must cover the six output coefficients");
    if (active && lane < 6) {
        // w-power order of the program outputs -> tower slots c0.c0, c1.c0, c0.c1, c1.c1, c0.c2, c1.c2
        Fp2 v;
        if (trivial) v = (lane == 0) ? fp2_one() : fp2_zero();
        else { const int outs[6] = {P::miller_out[0], P::miller_out[1], P::miller_out[2], P::miller_out[3], P::miller_out[4], P::miller_out[5]}; v = rf.load(uint32_t(outs[lane])); }
        Fp2* dst = reinterpret_cast<Fp2*>(f + i);
        const int tower_pos[6] = {0, 3, 1, 4, 2, 5};  // Fp12 memory order: c0.{c0,c1,c2}, c1.{c0,c1,c2}
        dst[tower_pos[lane]] = v;
    }
}

// one team per tuple (two Miller values per tuple: pairs pair_off[t], pair_off[t]+1)
template <int TEAM>
__global__ void __launch_bounds__(128) k_vm_final(const uint32_t* __restrict__ code, const Fp2* __restrict__ consts,
                                                  const Fp12* __restrict__ f, const uint32_t* __restrict__ pair_off,
                                                  const int32_t* __restrict__ pk_code, const uint32_t* __restrict__ flags,
                                                  const int32_t* __restrict__ sig_code, uint32_t n_tuples,
                                                  int32_t* __restrict__ out_codes) {
    typedef VmProg<TEAM> P;
    extern __shared__ uint32_t smem[];
    const uint32_t team_in_block = threadIdx.x / TEAM, lane = threadIdx.x % TEAM;
    const uint32_t t = blockIdx.x * (blockDim.x / TEAM) + team_in_block;
    VmRfStrided rf{smem + team_in_block * (P::final_slots * kVmSlotWords)};
    int32_t code_out = BLS_SUCCESS;
    bool active = false;
    if (t < n_tuples) {
        if (pk_code[t] != BLS_SUCCESS) code_out = pk_code[t];
        else if (sig_code[t] > 0) code_out = sig_code[t];
        else if (flags[t] != 0 || sig_code[t] == SIG_NOT_IN_GROUP) code_out = BLS_VERIFY_FAIL;
        else active = true;
    }
    if (active) {
        const int tower_pos[6] = {0, 3, 1, 4, 2, 5};
        for (uint32_t k = lane; k < 12; k += TEAM) {
            const Fp2* src = reinterpret_cast<const Fp2*>(f + pair_off[t] + k / 6);
            rf.store(k, src[tower_pos[k % 6]]);
        }
    }
    __syncwarp();
    vm_run<TEAM>(code, P::final_rounds, consts, rf, lane, active);
    if (t < n_tuples && lane == 0) {
        if (active) {
            const int outs[6] = {P::final_out[0], P::final_out[1], P::final_out[2], P::final_out[3], P::final_out[4], P::final_out[5]};
            bool one = fp2_eq(rf.load(uint32_t(outs[0])), fp2_one());
#pragma unroll 1
            for (int k = 1; k < 6; k++) one = one && fp2_is_zero(rf.load(uint32_t(outs[k])));
            code_out = one ? BLS_SUCCESS : BLS_VERIFY_FAIL;
        }
        out_codes[t] = code_out;
    }
}

}  // namespace

// device copies of the two program pairs: [0] teams of 8 lanes, [1] teams of 16
static uint32_t* g_d_miller_code[2] = {nullptr, nullptr};
static uint32_t* g_d_final_code[2] = {nullptr, nullptr};
static Fp2* g_d_consts = nullptr;
// Batches with at most this many teams' worth of work run on 16-lane teams: they cannot fill the machine anyway, so the
// shorter critical path (1 918 / 3 094 rounds instead of 2 493 / 3 842) wins; above it the 8-lane programs' higher
// throughput does.  B200_VM_TEAM16_MAX overrides (0: never).
static uint32_t g_team16_max = 2048;

template <int TEAM>
static int vm_upload(int slot, cudaStream_t st) {
    typedef VmProg<TEAM> P;
    if (cudaMalloc(&g_d_miller_code[slot], P::miller_code_bytes) != cudaSuccess) return 1;
    if (cudaMalloc(&g_d_final_code[slot], P::final_code_bytes) != cudaSuccess) return 1;
    cudaMemcpyAsync(g_d_miller_code[slot], P::miller_code(), P::miller_code_bytes, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(g_d_final_code[slot], P::final_code(), P::final_code_bytes, cudaMemcpyHostToDevice, st);
    cudaFuncSetAttribute(k_vm_miller<TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 / TEAM) * P::miller_slots * kVmSlotWords * 4);
    cudaFuncSetAttribute(k_vm_final<TEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (128 / TEAM) * P::final_slots * kVmSlotWords * 4);
    return 0;
}

int vm_init(void* stream) {
    if (g_d_consts) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (const char* v = getenv("B200_VM_TEAM16_MAX")) g_team16_max = uint32_t(atol(v));
    uint32_t* d_plain = nullptr;
    if (vm_upload<8>(0, st) || vm_upload<16>(1, st)) return 1;
    if (cudaMalloc(&g_d_consts, sizeof(Fp2) * kVmConsts) != cudaSuccess) return 1;
    if (cudaMalloc(&d_plain, sizeof(h_vm_consts)) != cudaSuccess) return 1;
    cudaMemcpyAsync(d_plain, h_vm_consts, sizeof(h_vm_consts), cudaMemcpyHostToDevice, st);
    k_vm_consts<<<1, 32, 0, st>>>(d_plain, g_d_consts);
    if (cudaStreamSynchronize(st) != cudaSuccess) return 1;
    cudaFree(d_plain);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

template <int TEAM>
static void launch_vm_miller_t(int slot, const G1Pre* g1, const uint32_t* g1_idx, const G2Aff* g2, const uint32_t* g2_idx,
                               const uint32_t* pair_tuple, const int32_t* pk_code, const uint32_t* flags, const int32_t* sig_code,
                               uint32_t n_pairs, Fp12* f, cudaStream_t st) {
    const int threads = 128, teams = threads / TEAM;
    const size_t smem = size_t(teams) * VmProg<TEAM>::miller_slots * kVmSlotWords * 4;
    k_vm_miller<TEAM><<<(n_pairs + teams - 1) / teams, threads, smem, st>>>(
        g_d_miller_code[slot], g_d_consts, g1, g1_idx, g2, g2_idx, pair_tuple, pk_code, flags, sig_code, n_pairs, f);
}
template <int TEAM>
static void launch_vm_final_t(int slot, const Fp12* f, const uint32_t* pair_off, const int32_t* pk_code, const uint32_t* flags,
                              const int32_t* sig_code, uint32_t n_tuples, int32_t* out_codes, cudaStream_t st) {
    const int threads = 128, teams = threads / TEAM;
    const size_t smem = size_t(teams) * VmProg<TEAM>::final_slots * kVmSlotWords * 4;
    k_vm_final<TEAM><<<(n_tuples + teams - 1) / teams, threads, smem, st>>>(
        g_d_final_code[slot], g_d_consts, f, pair_off, pk_code, flags, sig_code, n_tuples, out_codes);
}

void launch_vm_miller(const G1Pre* g1, const uint32_t* g1_idx, const G2Aff* g2, const uint32_t* g2_idx,
                      const uint32_t* pair_tuple, const int32_t* pk_code, const uint32_t* flags, const int32_t* sig_code,
                      uint32_t n_pairs, Fp12* f, void* stream) {
    if (!n_pairs) return;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_pairs <= g_team16_max) launch_vm_miller_t<16>(1, g1, g1_idx, g2, g2_idx, pair_tuple, pk_code, flags, sig_code, n_pairs, f, st);
    else launch_vm_miller_t<8>(0, g1, g1_idx, g2, g2_idx, pair_tuple, pk_code, flags, sig_code, n_pairs, f, st);
}
void launch_vm_final(const Fp12* f, const uint32_t* pair_off, const int32_t* pk_code, const uint32_t* flags,
                     const int32_t* sig_code, uint32_t n_tuples, int32_t* out_codes, void* stream) {
    if (!n_tuples) return;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_tuples <= g_team16_max) launch_vm_final_t<16>(1, f, pair_off, pk_code, flags, sig_code, n_tuples, out_codes, st);
    else launch_vm_final_t<8>(0, f, pair_off, pk_code, flags, sig_code, n_tuples, out_codes, st);
}

}  // namespace b200
