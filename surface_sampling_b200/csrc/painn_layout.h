// Packed PaiNN weight layout (floats) for one model.  Python mirror:
// surface_sampling_b200/engine.py::PAINN_LAYOUT (a test checks both totals agree).
// "T" = transposed to [in][out] so it is the row-major B operand of  C[M,N] = A[M,K] . B[K,N];
// the un-transposed copies are the B operands of the backward (input-gradient) GEMMs.
#pragma once

namespace painn {
constexpr int F = 128;      // feat_dim
constexpr int F3 = 384;
constexpr int NRBF = 20;
constexpr int NCONV = 3;
constexpr int NEMB = 100;
constexpr int FH = 64;      // readout hidden

// per-layer block
constexpr long long L_W1T = 0;                       // [128][128]
constexpr long long L_B1 = L_W1T + F * F;            // [128]
constexpr long long L_W2T = L_B1 + F;                // [128][384]
constexpr long long L_B2 = L_W2T + F * F3;           // [384]
constexpr long long L_WDT = L_B2 + F3;               // [20][384]
constexpr long long L_BD = L_WDT + NRBF * F3;        // [384]
constexpr long long L_UVT = L_BD + F3;               // [128][256]  (U^T | V^T)
constexpr long long L_W3T = L_UVT + F * 2 * F;       // [256][128]
constexpr long long L_B3 = L_W3T + 2 * F * F;        // [128]
constexpr long long L_W4T = L_B3 + F;                // [128][384]
constexpr long long L_B4 = L_W4T + F * F3;           // [384]
constexpr long long L_W1 = L_B4 + F3;                // [128][128]
constexpr long long L_W2 = L_W1 + F * F;             // [384][128]
constexpr long long L_UV = L_W2 + F3 * F;            // [256][128]  (U rows ; V rows)
constexpr long long L_W3 = L_UV + 2 * F * F;         // [128][256]
constexpr long long L_W4 = L_W3 + F * 2 * F;         // [384][128]
constexpr long long L_SIZE = L_W4 + F3 * F;

constexpr long long W_EMBED = 0;                     // [100][128]
constexpr long long W_LAYER0 = W_EMBED + NEMB * F;
constexpr long long W_READ = W_LAYER0 + NCONV * L_SIZE;
constexpr long long R_W5T = W_READ;                  // [128][64]
constexpr long long R_B5 = R_W5T + F * FH;           // [64]
constexpr long long R_W6 = R_B5 + FH;                // [64]
constexpr long long R_B6 = R_W6 + FH;                // [1] (+3 pad)
constexpr long long R_W5 = R_B6 + 4;                 // [64][128]
constexpr long long W_TOTAL = R_W5 + FH * F;
// packed block per model = [exact fp32 | TF32 part (hi) | remainder (lo = w - hi)], each W_TOTAL floats;
// the hi/lo copies feed the 3xTF32 tcgen05 GEMM (gemm_tc.cuh)
constexpr long long W_STRIDE = 3 * W_TOTAL;
}  // namespace painn
