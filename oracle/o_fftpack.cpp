/* ORACLE (test infrastructure) — literal restatement of the reachable part of the vendored
 * real FFTPACK: reference source/fftpack.f90.  Only radix 2/3/4 are reachable for
 * n = 96 (factors 2,4,4,3) and n = 144 (4,4,3,3); radix 5 / generic abort.
 * All perturbed single-precision constants are kept:
 *   tpi = 8.*atan(1.) (fftpack.f90:39), taui = .5*sqrt(3.) (:269,:787),
 *   sqrt2 = sqrt(2.) (:341), hsqt2 = .5*sqrt(2.) (:857).
 */
#include "oracle.h"

namespace orc {

/* fftpack.f90:1-67 */
void rffti1(int n, double* wa /*1-based*/, int* ifac /*1-based*/) {
    static const int ntryh[5] = {0, 4, 2, 3, 5};
    int nl = n, nf = 0, j = 0, ntry = 0;
    for (;;) {
        j = j + 1;
        if (j <= 4) ntry = ntryh[j]; else ntry = ntry + 2;
        for (;;) {
            int nq = nl / ntry;
            int nr = nl - ntry * nq;
            if (nr != 0) break;
            nf = nf + 1;
            ifac[nf + 2] = ntry;
            nl = nq;
            if (ntry == 2 && nf != 1) {
                for (int i = 2; i <= nf; i++) {
                    int ib = nf - i + 2;
                    ifac[ib + 2] = ifac[ib + 1];
                }
                ifac[3] = 2;
            }
            if (nl == 1) goto done;
        }
    }
done:
    ifac[1] = n;
    ifac[2] = nf;
    float tpi_f = 8.f * atanf(1.f);            /* :39 real32 */
    double tpi = (double)tpi_f;
    double argh = tpi / n;                     /* :40 real64 quotient */
    int is = 0;
    int nfm1 = nf - 1;
    int l1 = 1;
    if (nfm1 == 0) return;
    for (int k1 = 1; k1 <= nfm1; k1++) {
        int ip = ifac[k1 + 2];
        int ld = 0;
        int l2 = l1 * ip;
        int ido = n / l2;
        int ipm = ip - 1;
        for (int jj = 1; jj <= ipm; jj++) {
            ld = ld + l1;
            int i = is;
            double argld = ld * argh;
            double fi = 0.;
            for (int ii = 3; ii <= ido; ii += 2) {
                i = i + 2;
                fi = fi + 1.;
                double arg = fi * argld;
                wa[i - 1] = cos(arg);
                wa[i] = sin(arg);
            }
            is = is + ido;
        }
        l1 = l2;
    }
}

#define CC(i, j, k) cc[((i)-1) + ido * (((j)-1) + IP * ((k)-1))]
#define CH(i, k, j) ch[((i)-1) + ido * (((k)-1) + l1 * ((j)-1))]

/* fftpack.f90:204-254 */
static void radb2(int ido, int l1, const double* cc, double* ch, const double* wa1 /*1-based*/) {
    const int IP = 2;
    for (int k = 1; k <= l1; k++) {
        CH(1, k, 1) = CC(1, 1, k) + CC(ido, 2, k);
        CH(1, k, 2) = CC(1, 1, k) - CC(ido, 2, k);
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                CH(i - 1, k, 1) = CC(i - 1, 1, k) + CC(ic - 1, 2, k);
                double tr2 = CC(i - 1, 1, k) - CC(ic - 1, 2, k);
                CH(i, k, 1) = CC(i, 1, k) - CC(ic, 2, k);
                double ti2 = CC(i, 1, k) + CC(ic, 2, k);
                CH(i - 1, k, 2) = wa1[i - 2] * tr2 - wa1[i - 1] * ti2;
                CH(i, k, 2) = wa1[i - 2] * ti2 + wa1[i - 1] * tr2;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        CH(ido, k, 1) = CC(ido, 1, k) + CC(ido, 1, k);
        CH(ido, k, 2) = -(CC(1, 2, k) + CC(1, 2, k));
    }
}

/* fftpack.f90:256-326 */
static void radb3(int ido, int l1, const double* cc, double* ch, const double* wa1, const double* wa2) {
    const int IP = 3;
    const double taur = -.5;
    const double taui = (double)(.5f * sqrtf(3.f));
    for (int k = 1; k <= l1; k++) {
        double tr2 = CC(ido, 2, k) + CC(ido, 2, k);
        double cr2 = CC(1, 1, k) + taur * tr2;
        CH(1, k, 1) = CC(1, 1, k) + tr2;
        double ci3 = taui * (CC(1, 3, k) + CC(1, 3, k));
        CH(1, k, 2) = cr2 - ci3;
        CH(1, k, 3) = cr2 + ci3;
    }
    if (ido == 1) return;
    int idp2 = ido + 2;
    for (int k = 1; k <= l1; k++)
        for (int i = 3; i <= ido; i += 2) {
            int ic = idp2 - i;
            double tr2 = CC(i - 1, 3, k) + CC(ic - 1, 2, k);
            double cr2 = CC(i - 1, 1, k) + taur * tr2;
            CH(i - 1, k, 1) = CC(i - 1, 1, k) + tr2;
            double ti2 = CC(i, 3, k) - CC(ic, 2, k);
            double ci2 = CC(i, 1, k) + taur * ti2;
            CH(i, k, 1) = CC(i, 1, k) + ti2;
            double cr3 = taui * (CC(i - 1, 3, k) - CC(ic - 1, 2, k));
            double ci3 = taui * (CC(i, 3, k) + CC(ic, 2, k));
            double dr2 = cr2 - ci3;
            double dr3 = cr2 + ci3;
            double di2 = ci2 + cr3;
            double di3 = ci2 - cr3;
            CH(i - 1, k, 2) = wa1[i - 2] * dr2 - wa1[i - 1] * di2;
            CH(i, k, 2) = wa1[i - 2] * di2 + wa1[i - 1] * dr2;
            CH(i - 1, k, 3) = wa2[i - 2] * dr3 - wa2[i - 1] * di3;
            CH(i, k, 3) = wa2[i - 2] * di3 + wa2[i - 1] * dr3;
        }
}

/* fftpack.f90:328-424 */
static void radb4(int ido, int l1, const double* cc, double* ch, const double* wa1, const double* wa2, const double* wa3) {
    const int IP = 4;
    const double sqrt2 = (double)sqrtf(2.f);
    for (int k = 1; k <= l1; k++) {
        double tr1 = CC(1, 1, k) - CC(ido, 4, k);
        double tr2 = CC(1, 1, k) + CC(ido, 4, k);
        double tr3 = CC(ido, 2, k) + CC(ido, 2, k);
        double tr4 = CC(1, 3, k) + CC(1, 3, k);
        CH(1, k, 1) = tr2 + tr3;
        CH(1, k, 2) = tr1 - tr4;
        CH(1, k, 3) = tr2 - tr3;
        CH(1, k, 4) = tr1 + tr4;
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                double ti1 = CC(i, 1, k) + CC(ic, 4, k);
                double ti2 = CC(i, 1, k) - CC(ic, 4, k);
                double ti3 = CC(i, 3, k) - CC(ic, 2, k);
                double tr4 = CC(i, 3, k) + CC(ic, 2, k);
                double tr1 = CC(i - 1, 1, k) - CC(ic - 1, 4, k);
                double tr2 = CC(i - 1, 1, k) + CC(ic - 1, 4, k);
                double ti4 = CC(i - 1, 3, k) - CC(ic - 1, 2, k);
                double tr3 = CC(i - 1, 3, k) + CC(ic - 1, 2, k);
                CH(i - 1, k, 1) = tr2 + tr3;
                double cr3 = tr2 - tr3;
                CH(i, k, 1) = ti2 + ti3;
                double ci3 = ti2 - ti3;
                double cr2 = tr1 - tr4;
                double cr4 = tr1 + tr4;
                double ci2 = ti1 + ti4;
                double ci4 = ti1 - ti4;
                CH(i - 1, k, 2) = wa1[i - 2] * cr2 - wa1[i - 1] * ci2;
                CH(i, k, 2) = wa1[i - 2] * ci2 + wa1[i - 1] * cr2;
                CH(i - 1, k, 3) = wa2[i - 2] * cr3 - wa2[i - 1] * ci3;
                CH(i, k, 3) = wa2[i - 2] * ci3 + wa2[i - 1] * cr3;
                CH(i - 1, k, 4) = wa3[i - 2] * cr4 - wa3[i - 1] * ci4;
                CH(i, k, 4) = wa3[i - 2] * ci4 + wa3[i - 1] * cr4;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        double ti1 = CC(1, 2, k) + CC(1, 4, k);
        double ti2 = CC(1, 4, k) - CC(1, 2, k);
        double tr1 = CC(ido, 1, k) - CC(ido, 3, k);
        double tr2 = CC(ido, 1, k) + CC(ido, 3, k);
        CH(ido, k, 1) = tr2 + tr2;
        CH(ido, k, 2) = sqrt2 * (tr1 - ti1);
        CH(ido, k, 3) = ti2 + ti2;
        CH(ido, k, 4) = -sqrt2 * (tr1 + ti1);
    }
}
#undef CC
#undef CH

/* forward passes: cc(ido,l1,ip) -> ch(ido,ip,l1) */
#define CC(i, k, j) cc[((i)-1) + ido * (((k)-1) + l1 * ((j)-1))]
#define CH(i, j, k) ch[((i)-1) + ido * (((j)-1) + IP * ((k)-1))]

/* fftpack.f90:722-772 */
static void radf2(int ido, int l1, const double* cc, double* ch, const double* wa1) {
    const int IP = 2;
    for (int k = 1; k <= l1; k++) {
        CH(1, 1, k) = CC(1, k, 1) + CC(1, k, 2);
        CH(ido, 2, k) = CC(1, k, 1) - CC(1, k, 2);
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                double tr2 = wa1[i - 2] * CC(i - 1, k, 2) + wa1[i - 1] * CC(i, k, 2);
                double ti2 = wa1[i - 2] * CC(i, k, 2) - wa1[i - 1] * CC(i - 1, k, 2);
                CH(i, 1, k) = CC(i, k, 1) + ti2;
                CH(ic, 2, k) = ti2 - CC(i, k, 1);
                CH(i - 1, 1, k) = CC(i - 1, k, 1) + tr2;
                CH(ic - 1, 2, k) = CC(i - 1, k, 1) - tr2;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        CH(1, 2, k) = -CC(ido, k, 2);
        CH(ido, 1, k) = CC(ido, k, 1);
    }
}

/* fftpack.f90:774-842 */
static void radf3(int ido, int l1, const double* cc, double* ch, const double* wa1, const double* wa2) {
    const int IP = 3;
    const double taur = -.5;
    const double taui = (double)(.5f * sqrtf(3.f));
    for (int k = 1; k <= l1; k++) {
        double cr2 = CC(1, k, 2) + CC(1, k, 3);
        CH(1, 1, k) = CC(1, k, 1) + cr2;
        CH(1, 3, k) = taui * (CC(1, k, 3) - CC(1, k, 2));
        CH(ido, 2, k) = CC(1, k, 1) + taur * cr2;
    }
    if (ido == 1) return;
    int idp2 = ido + 2;
    for (int k = 1; k <= l1; k++)
        for (int i = 3; i <= ido; i += 2) {
            int ic = idp2 - i;
            double dr2 = wa1[i - 2] * CC(i - 1, k, 2) + wa1[i - 1] * CC(i, k, 2);
            double di2 = wa1[i - 2] * CC(i, k, 2) - wa1[i - 1] * CC(i - 1, k, 2);
            double dr3 = wa2[i - 2] * CC(i - 1, k, 3) + wa2[i - 1] * CC(i, k, 3);
            double di3 = wa2[i - 2] * CC(i, k, 3) - wa2[i - 1] * CC(i - 1, k, 3);
            double cr2 = dr2 + dr3;
            double ci2 = di2 + di3;
            CH(i - 1, 1, k) = CC(i - 1, k, 1) + cr2;
            CH(i, 1, k) = CC(i, k, 1) + ci2;
            double tr2 = CC(i - 1, k, 1) + taur * cr2;
            double ti2 = CC(i, k, 1) + taur * ci2;
            double tr3 = taui * (di2 - di3);
            double ti3 = taui * (dr3 - dr2);
            CH(i - 1, 3, k) = tr2 + tr3;
            CH(ic - 1, 2, k) = tr2 - tr3;
            CH(i, 3, k) = ti2 + ti3;
            CH(ic, 2, k) = ti3 - ti2;
        }
}

/* fftpack.f90:844-936 */
static void radf4(int ido, int l1, const double* cc, double* ch, const double* wa1, const double* wa2, const double* wa3) {
    const int IP = 4;
    const double hsqt2 = (double)(.5f * sqrtf(2.f));
    for (int k = 1; k <= l1; k++) {
        double tr1 = CC(1, k, 2) + CC(1, k, 4);
        double tr2 = CC(1, k, 1) + CC(1, k, 3);
        CH(1, 1, k) = tr1 + tr2;
        CH(ido, 4, k) = tr2 - tr1;
        CH(ido, 2, k) = CC(1, k, 1) - CC(1, k, 3);
        CH(1, 3, k) = CC(1, k, 4) - CC(1, k, 2);
    }
    if (ido < 2) return;
    if (ido > 2) {
        int idp2 = ido + 2;
        for (int k = 1; k <= l1; k++)
            for (int i = 3; i <= ido; i += 2) {
                int ic = idp2 - i;
                double cr2 = wa1[i - 2] * CC(i - 1, k, 2) + wa1[i - 1] * CC(i, k, 2);
                double ci2 = wa1[i - 2] * CC(i, k, 2) - wa1[i - 1] * CC(i - 1, k, 2);
                double cr3 = wa2[i - 2] * CC(i - 1, k, 3) + wa2[i - 1] * CC(i, k, 3);
                double ci3 = wa2[i - 2] * CC(i, k, 3) - wa2[i - 1] * CC(i - 1, k, 3);
                double cr4 = wa3[i - 2] * CC(i - 1, k, 4) + wa3[i - 1] * CC(i, k, 4);
                double ci4 = wa3[i - 2] * CC(i, k, 4) - wa3[i - 1] * CC(i - 1, k, 4);
                double tr1 = cr2 + cr4;
                double tr4 = cr4 - cr2;
                double ti1 = ci2 + ci4;
                double ti4 = ci2 - ci4;
                double ti2 = CC(i, k, 1) + ci3;
                double ti3 = CC(i, k, 1) - ci3;
                double tr2 = CC(i - 1, k, 1) + cr3;
                double tr3 = CC(i - 1, k, 1) - cr3;
                CH(i - 1, 1, k) = tr1 + tr2;
                CH(ic - 1, 4, k) = tr2 - tr1;
                CH(i, 1, k) = ti1 + ti2;
                CH(ic, 4, k) = ti1 - ti2;
                CH(i - 1, 3, k) = ti4 + tr3;
                CH(ic - 1, 2, k) = tr3 - ti4;
                CH(i, 3, k) = tr4 + ti3;
                CH(ic, 2, k) = tr4 - ti3;
            }
        if (ido % 2 == 1) return;
    }
    for (int k = 1; k <= l1; k++) {
        double ti1 = -hsqt2 * (CC(ido, k, 2) + CC(ido, k, 4));
        double tr1 = hsqt2 * (CC(ido, k, 2) - CC(ido, k, 4));
        CH(ido, 1, k) = tr1 + CC(ido, k, 1);
        CH(ido, 3, k) = CC(ido, k, 1) - tr1;
        CH(1, 2, k) = ti1 - CC(ido, k, 3);
        CH(1, 4, k) = ti1 + CC(ido, k, 3);
    }
}
#undef CC
#undef CH

static void unreachable_radix(int ip) {
    fprintf(stderr, "oracle fftpack: radix %d not reachable for n=96/144 (fftpack.f90)\n", ip);
    abort();
}

/* fftpack.f90:69-134; c, ch, wa 1-based (pass pointer-1) */
void rfftb1(int n, double* c, double* ch, const double* wa, const int* ifac) {
    int nf = ifac[2];
    int na = 0;
    int l1 = 1;
    int iw = 1;
    for (int k1 = 1; k1 <= nf; k1++) {
        int ip = ifac[k1 + 2];
        int l2 = ip * l1;
        int ido = n / l2;
        double* in = (na == 0) ? c : ch;
        double* out = (na == 0) ? ch : c;
        if (ip == 4) {
            int ix2 = iw + ido, ix3 = ix2 + ido;
            radb4(ido, l1, in + 1, out + 1, wa + iw - 1, wa + ix2 - 1, wa + ix3 - 1);
            na = 1 - na;
        } else if (ip == 2) {
            radb2(ido, l1, in + 1, out + 1, wa + iw - 1);
            na = 1 - na;
        } else if (ip == 3) {
            int ix2 = iw + ido;
            radb3(ido, l1, in + 1, out + 1, wa + iw - 1, wa + ix2 - 1);
            na = 1 - na;
        } else {
            unreachable_radix(ip);
        }
        l1 = l2;
        iw = iw + (ip - 1) * ido;
    }
    if (na == 0) return;
    for (int i = 1; i <= n; i++) c[i] = ch[i];
}

/* fftpack.f90:136-202 */
void rfftf1(int n, double* c, double* ch, const double* wa, const int* ifac) {
    int nf = ifac[2];
    int na = 1;
    int l2 = n;
    int iw = n;
    for (int k1 = 1; k1 <= nf; k1++) {
        int kh = nf - k1;
        int ip = ifac[kh + 3];
        int l1 = l2 / ip;
        int ido = n / l2;
        iw = iw - (ip - 1) * ido;
        na = 1 - na;
        double* in = (na == 0) ? c : ch;
        double* out = (na == 0) ? ch : c;
        if (ip == 4) {
            int ix2 = iw + ido, ix3 = ix2 + ido;
            radf4(ido, l1, in + 1, out + 1, wa + iw - 1, wa + ix2 - 1, wa + ix3 - 1);
        } else if (ip == 2) {
            radf2(ido, l1, in + 1, out + 1, wa + iw - 1);
        } else if (ip == 3) {
            int ix2 = iw + ido;
            radf3(ido, l1, in + 1, out + 1, wa + iw - 1, wa + ix2 - 1);
        } else {
            unreachable_radix(ip);
        }
        l2 = l1;
    }
    if (na == 1) return;
    for (int i = 1; i <= n; i++) c[i] = ch[i];
}

}  // namespace orc
