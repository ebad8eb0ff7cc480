// fs_perturb_loop.cuh -- the plain perturbation loop with rebasing (LAKernel.cuh:130-236), generic over the
// numeric policy, plus the hand-scheduled HDRx32 version.
//
// Why a special HDRx32 loop: ncu on the generic loop (profiles/r1_lav2_v0_summary.md) shows the half-rate
// ALU pipe 91 % busy with the selects/compares of the float+exponent alignment while the FMA pipe idles at
// 22 %.  The fast step below computes the SAME roundings with a different instruction mix:
//   * alignment of (a.m, a.e) + (b.m, b.e) without selects: both operands are scaled, one of the two
//     multipliers is exactly 1:  m = fma(b.m, 2^min(-d,0), a.m * 2^min(d,0)),  d = a.e - b.e.
//     Each multiplier's exponent field clamp(127 -/+ d, 0, 127) is ONE DPX instruction (VIADDMNMX.RELU),
//     and lands on 0.0f exactly when the reference's getMultiplierNeg returns 0 (|d| >= 127).
//   * integer adds / shifts are issued as IMAD with opaque multipliers so they run on the FMA pipe;
//   * Reduce() is one PRMT + one 3-input add + one LOP3;
//   * the reference's zero-mantissa special cases (add/sub reset the exponent, Reduce is a no-op) are not
//     evaluated per operation: one product of the step's mantissas is tested, and a step that touched an exact
//     zero is recomputed by the generic, reference-shaped code.  Results are bit-identical by construction in
//     both branches; tests/test_gpu_parity.py checks that against the reference kernels.
#pragma once
#include "fs_num.cuh"

namespace fs {

template <class Num> struct OrbitIO; // fs_lav2.cuh

// ---- generic loop (any numeric policy) --------------------------------------------------------------------
template <class Num, class IterT, bool Count> struct PerturbLoop {
    using Real = typename Num::Real;
    FS_D static void run(const void *orbit, IterT orbit_count, IterT n_iterations, Real dcX, Real dcY, Real &dX,
                         Real &dY, IterT &RefIteration, IterT &iter, unsigned long long &steps) {
        Real zx, zy;
        OrbitIO<Num>::load(orbit, RefIteration, zx, zy);
        const IterT last = orbit_count - 1;
        for (;;) {
            Num::perturb(dX, dY, zx, zy, dcX, dcY);
            ++RefIteration;
            OrbitIO<Num>::load(orbit, RefIteration, zx, zy);
            const Real tX = add(zx, dX);
            const Real tY = add(zy, dY);
            const Real n2 = Num::norm2(tX, tY);
            if (Count) steps++;
            if (lt_bailout(n2) && iter < n_iterations) {
                const Real d2 = Num::norm2(dX, dY);
                if (lt_pr(n2, d2) || RefIteration >= last) {
                    dX = tX;
                    dY = tY;
                    RefIteration = 0;
                    OrbitIO<Num>::load(orbit, 0, zx, zy);
                }
                ++iter;
            } else {
                break;
            }
        }
    }
};

// ---- HDRx32: select-free step ------------------------------------------------------------------------------
struct Hdr32Fast {
    // constants the compiler must not see through (keeps IMAD on the FMA pipe instead of IADD3/SHF on the ALU pipe)
    int M1;  // -1
    int K23; // 1 << 23
    int K1;  // 1
    FS_D void init() {
        asm volatile("mov.s32 %0, -1;" : "=r"(M1));
        asm volatile("mov.s32 %0, 8388608;" : "=r"(K23));
        asm volatile("mov.s32 %0, 1;" : "=r"(K1));
    }
    // (a.m, a.e) + (b.m, b.e): mantissa returned, exponent max(a.e, b.e) in E
    FS_D float align(float am, int ae, float bm, int be, int &E) const {
        const int d = be * M1 + ae;  // a.e - b.e
        const int nd = ae * M1 + be; // b.e - a.e
        const int fa = __viaddmin_s32_relu(d, 127, 127);  // clamp(127 + d, 0, 127): field of 2^min(d,0)
        const int fb = __viaddmin_s32_relu(nd, 127, 127); // clamp(127 - d, 0, 127): field of 2^min(-d,0)
        const float ma = __int_as_float(fa * K23);
        const float mb = __int_as_float(fb * K23);
        E = max(ae, be);
        return __fmaf_rn(bm, mb, am * ma);
    }
    // Reduce() of a non-zero mantissa (HDRFloat.h:432-448)
    FS_D void reduce_nz(float &m, int &e) const {
        const unsigned b = __float_as_uint(m);
        const unsigned fe = __byte_perm(b + b, 0u, 0x4443); // exponent field: byte 3 of (bits << 1)
        e = e + (int)fe - 127;
        m = __uint_as_float((b & 0x807fffffu) | 0x3f800000u);
    }
    FS_D void reduce_pos(float &m, int &e) const { // mantissa known positive
        const unsigned b = __float_as_uint(m);
        e = e + (int)(b >> 23) - 127;
        m = __uint_as_float((b & 0x007fffffu) | 0x3f800000u);
    }
};

template <class IterT, bool Count> struct PerturbLoop<NumHdr<float>, IterT, Count> {
    using Num = NumHdr<float>;
    using Real = Hdr<float>;
    FS_D static void run(const void *orbit, IterT orbit_count, IterT n_iterations, Real dcX, Real dcY, Real &dXio,
                         Real &dYio, IterT &RefIteration, IterT &iter, unsigned long long &steps) {
        Hdr32Fast F;
        F.init();
        const uint4 *__restrict__ orb = reinterpret_cast<const uint4 *>(orbit);
        const IterT last = orbit_count - 1;
        float dxm = dXio.m, dym = dYio.m;
        int dxe = dXio.e, dye = dYio.e;
        const float cxm = dcX.m, cym = dcY.m;
        const int cxe = dcX.e, cye = dcY.e;
        uint4 z = __ldg(orb + RefIteration); // {x.m, x.e, y.e, y.m}
        for (;;) {
            ++RefIteration;
            const uint4 zn = __ldg(orb + RefIteration);
            const float zxm = __uint_as_float(z.x), zym = __uint_as_float(z.w);
            const int zxe1 = (int)z.y + 1, zye1 = (int)z.z + 1; // 2*Z: exponent + 1 (the clamp at MIN_BIG cannot trigger)
            // tempSum2 = 2Zx + dx ; tempSum1 = 2Zy + dy
            int s2e, s1e;
            const float s2m = F.align(zxm, zxe1, dxm, dxe, s2e);
            const float s1m = F.align(zym, zye1, dym, dye, s1e);
            // custom_perturb2, X: dx*s2 - dy*s1 + cx
            const float pXa = dxm * s2m, pXb = dym * s1m;
            int eX, nxe;
            const float sumX = F.align(pXa, dxe + s2e, -pXb, dye + s1e, eX);
            float nxm = F.align(sumX, eX, cxm, cxe, nxe);
            // Y: dx*s1 + dy*s2 + cy
            const float pYa = dxm * s1m, pYb = dym * s2m;
            int eY, nye;
            const float sumY = F.align(pYa, dxe + s1e, pYb, dye + s2e, eY);
            float nym = F.align(sumY, eY, cym, cye, nye);
            const float nz_guard = nxm * nym; // raw (unreduced) results: zero here needs the reference's special cases
            F.reduce_nz(nxm, nxe);
            F.reduce_nz(nym, nye);
            // z = Z' + d'
            const float wxm = __uint_as_float(zn.x), wym = __uint_as_float(zn.w);
            int txe, tye;
            float txm = F.align(wxm, (int)zn.y, nxm, nxe, txe);
            float tym = F.align(wym, (int)zn.z, nym, nye, tye);
            // |z|^2 and |d'|^2
            const float sqx = txm * txm, sqy = tym * tym, sdx = nxm * nxm, sdy = nym * nym;
            int n2e, d2e;
            float n2m = F.align(sqx, txe + txe, sqy, tye + tye, n2e);
            float d2m = F.align(sdx, nxe + nxe, sdy, nye + nye, d2e);
            F.reduce_pos(n2m, n2e);
            F.reduce_pos(d2m, d2e);
            bool below = n2e <= 1; // reduced, non-zero: "< 256" can never fail at exponent 1 (HDRFloat.h:1169-1184)
            bool rebase_cmp = n2e < d2e || (n2e == d2e && n2m < d2m);

            // one test for every exact-zero special case of the reference (see file header)
            const float guard = ((pXa * pXb) * (sqx * sqy)) * nz_guard;
            if (guard == 0.0f) {
                Real dX{dxm, dxe}, dY{dym, dye};
                const Real zx{zxm, (int)z.y}, zy{zym, (int)z.z};
                Num::perturb(dX, dY, zx, zy, dcX, dcY);
                const Real tX = add(Real{wxm, (int)zn.y}, dX), tY = add(Real{wym, (int)zn.z}, dY);
                const Real n2 = Num::norm2(tX, tY), d2 = Num::norm2(dX, dY);
                nxm = dX.m; nxe = dX.e; nym = dY.m; nye = dY.e;
                txm = tX.m; txe = tX.e; tym = tY.m; tye = tY.e;
                below = lt_bailout(n2);
                rebase_cmp = lt_pr(n2, d2);
            }
            if (Count) steps++;
            if (!(below && iter < n_iterations)) {
                dxm = nxm; dxe = nxe; dym = nym; dye = nye;
                break;
            }
            ++iter;
            if (rebase_cmp || RefIteration >= last) {
                dxm = txm; dxe = txe; dym = tym; dye = tye;
                RefIteration = 0;
                z = __ldg(orb);
            } else {
                dxm = nxm; dxe = nxe; dym = nym; dye = nye;
                z = zn;
            }
        }
        dXio.m = dxm; dXio.e = dxe; dYio.m = dym; dYio.e = dye;
    }
};

} // namespace fs
