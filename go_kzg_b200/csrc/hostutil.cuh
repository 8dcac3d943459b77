// hostutil.cuh -- host-side helpers shared by the C-ABI translation units: ABI <-> internal
// encodings, GLV/wNAF recoding of fixed scalars, point (de)compression, Fr domain tables.
#pragma once
#include <string.h>
#include <vector>
#include "g1.cuh"

namespace b200 {

// ABI encodings (include/b200_kzg.h): Fr = 4 x u64 canonical LE == 8 x u32 LE;
// G1 = X,Y,Z each 6 x u64 canonical LE == 36 x u32, infinity <=> Z == 0.
inline Fr fr_load_canon(const uint64_t* p) { Fr r; memcpy(r.l, p, 32); return r; }
inline void fr_store_canon(uint64_t* p, const Fr& a) { memcpy(p, a.l, 32); }
inline Fr fr_from_abi_mont(const uint64_t* p) { return fe_to_mont(fr_load_canon(p)); }
inline void fr_to_abi_from_mont(uint64_t* p, const Fr& a) { fr_store_canon(p, fe_from_mont(a)); }
inline bool fr_canon_valid(const uint64_t* p) {   // bls/bignum_all.go:12-35 ValidFr (value < r)
    Fr a = fr_load_canon(p);
    for (int i = 7; i >= 0; i--) {
        if (a.l[i] < FrParams::mod(i)) return true;
        if (a.l[i] > FrParams::mod(i)) return false;
    }
    return false;
}
inline Fr fr_from_u64(uint64_t v) { Fr a = Fr::zero(); a.l[0] = (uint32_t)v; a.l[1] = (uint32_t)(v >> 32); return fe_to_mont(a); }

inline G1J g1_from_abi(const uint64_t* p) {
    G1J r;
    memcpy(r.x.l, p, 48); memcpy(r.y.l, p + 6, 48); memcpy(r.z.l, p + 12, 48);
    r.x = fe_to_mont(r.x); r.y = fe_to_mont(r.y); r.z = fe_to_mont(r.z);
    return r;
}
inline void g1_to_abi(uint64_t* p, const G1J& a) {
    Fp x = fe_from_mont(a.x), y = fe_from_mont(a.y), z = fe_from_mont(a.z);
    memcpy(p, x.l, 48); memcpy(p + 6, y.l, 48); memcpy(p + 12, z.l, 48);
}

// ---- fixed-scalar recoding: k = k1 + k2 z^2, both halves as signed digit strings ----------
// mode 1: width-5 NAF (odd digits in [-15, 15]); mode 0: fixed 4-bit signed windows.
inline void wnaf5_128(unsigned __int128 v, int8_t* out /* B200_WNAF_LEN */) {
    memset(out, 0, B200_WNAF_LEN);
    int i = 0;
    while (v != 0) {
        int d = 0;
        if (v & 1) {
            d = (int)(v & 31);
            if (d >= 16) d -= 32;
            if (d >= 0) v -= (unsigned)d; else v += (unsigned)(-d);
        }
        out[i++] = (int8_t)d;
        v >>= 1;
    }
}
inline void window4_128(unsigned __int128 v, int8_t* out /* B200_WNAF_LEN */) {
    memset(out, 0, B200_WNAF_LEN);
    unsigned carry = 0;
    for (int w = 0; w < 33; w++) {
        unsigned d = (w < 32 ? (unsigned)((v >> (4 * w)) & 15) : 0u) + carry;
        int dd;
        if (d > 8) { dd = (int)d - 16; carry = 1; } else { dd = (int)d; carry = 0; }
        out[4 * w] = (int8_t)dd;
    }
}
inline void glv_split(const Fr& k_canon, unsigned __int128& k1, unsigned __int128& k2) {
    constexpr uint32_t z2l[4] = B200_GLV_Z2;
    unsigned __int128 z2 = 0;
    for (int i = 3; i >= 0; i--) z2 = (z2 << 32) | z2l[i];
    unsigned __int128 rem = 0, quo = 0;
    for (int bit = 255; bit >= 0; bit--) {
        int top = (int)(rem >> 127);
        rem = (rem << 1) | ((k_canon.l[bit >> 5] >> (bit & 31)) & 1u);
        quo <<= 1;
        if (top || rem >= z2) { rem -= z2; quo |= 1; }
    }
    k1 = rem; k2 = quo;
}
inline void make_scalar_program(ScalarProgram* sp, const Fr& k_canon, int mode) {
    unsigned __int128 k1, k2;
    glv_split(k_canon, k1, k2);
    if (mode == 1) { wnaf5_128(k1, sp->d1); wnaf5_128(k2, sp->d2); }
    else { window4_128(k1, sp->d1); window4_128(k2, sp->d2); }
    int top = -1;
    for (int i = 0; i < B200_WNAF_LEN; i++) if (sp->d1[i] || sp->d2[i]) top = i;
    sp->top = (int16_t)top;
    bool one = k_canon.l[0] == 1;
    for (int i = 1; i < 8; i++) one = one && k_canon.l[i] == 0;
    sp->is_one = one ? 1 : 0;
    sp->mode = (int8_t)mode;
    sp->pad[0] = sp->pad[1] = 0;
}

// ---- compression (bls/bls_kilic.go:114-121; ZCash 48-byte form) -------------------------
inline void g1_affine_canon(const G1J& p, Fp& x, Fp& y) {
    Fp zi = fe_inv(p.z), zi2 = fe_sqr(zi);
    x = fe_from_mont(fe_mul(p.x, zi2));
    y = fe_from_mont(fe_mul(p.y, fe_mul(zi2, zi)));
}
inline bool fp_canon_gt_half(const Fp& y) {   // y > (p-1)/2
    constexpr uint32_t half[12] = B200_FP_HALF;
    for (int i = 11; i >= 0; i--) {
        if (y.l[i] > half[i]) return true;
        if (y.l[i] < half[i]) return false;
    }
    return false;
}
inline void g1_compress(uint8_t out[48], const G1J& p) {
    memset(out, 0, 48);
    if (p.is_inf()) { out[0] = 0xC0; return; }
    Fp x, y;
    g1_affine_canon(p, x, y);
    for (int i = 0; i < 48; i++) out[i] = (uint8_t)(x.l[(47 - i) >> 2] >> (((47 - i) & 3) * 8));
    out[0] |= 0x80;
    if (fp_canon_gt_half(y)) out[0] |= 0x20;
}
// 0 ok, 1 malformed flags / x >= p, 2 not on curve, 3 not in the prime-order subgroup (kilic's FromCompressed,
// reached through bls/bls_kilic.go:118-121, rejects all three).
inline int g1_decompress(G1J& p, const uint8_t in[48]) {
    if (!(in[0] & 0x80)) return 1;
    if (in[0] & 0x40) {
        for (int i = 1; i < 48; i++) if (in[i]) return 1;
        if (in[0] & 0x3F) return 1;
        p = G1J::infinity();
        return 0;
    }
    Fp x = Fp::zero();
    for (int i = 0; i < 48; i++) {
        uint8_t b = in[i];
        if (i == 0) b &= 0x1F;
        x.l[(47 - i) >> 2] |= (uint32_t)b << (((47 - i) & 3) * 8);
    }
    bool lt = false;
    for (int i = 11; i >= 0; i--) {
        if (x.l[i] < FpParams::mod(i)) { lt = true; break; }
        if (x.l[i] > FpParams::mod(i)) break;
    }
    if (!lt) return 1;
    Fp xm = fe_to_mont(x);
    Fp rhs = fe_add(fe_mul(fe_sqr(xm), xm), fp_const_four());
    constexpr uint32_t e[12] = B200_FP_SQRT_EXP;
    Fp y = fe_pow<FpParams, 12>(rhs, e);
    if (fe_sqr(y) != rhs) return 2;
    Fp yc = fe_from_mont(y);
    if (fp_canon_gt_half(yc) != !!(in[0] & 0x20)) y = fe_neg(y);
    p.x = xm; p.y = y; p.z = Fp::one();
    if (!g1_in_subgroup(p)) return 3;
    return 0;
}

// ---- Fr domain (fft.go:21-61) ------------------------------------------------------------
inline Fr fr_scale2_root_canon(unsigned k) {   // bls/globals.go:27-60
    static const uint32_t roots[32][8] = B200_FR_ROOTS_OF_UNITY;
    Fr r; memcpy(r.l, roots[k], 32); return r;
}

}  // namespace b200
