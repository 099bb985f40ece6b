/* ORACLE (test infrastructure) — restatement of physics.f90, humidity.f90, convection.f90,
 * large_scale_condensation.f90, shortwave_radiation.f90, longwave_radiation.f90,
 * surface_fluxes.f90, vertical_diffusion.f90, sppt.f90, mod_radcon.f90 and auxiliaries.f90
 * of the reference.  Loop nests and statement order follow the Fortran; un-suffixed
 * literals are real32 widened to real64.  x**2.0, x**3.0, x**4.0
 * are written x*x, (x*x)*x, (x*x)*(x*x) (what gfortran -Ofast emits for constant exponents). */
#include "oracle.h"

namespace orc {

#define GLOOP for (int j = 1; j <= il; j++) for (int i = 1; i <= ix; i++)
/* the reference's i-outer / j-inner nests (order matters nowhere: no cross-column dependence) */
#define IJLOOP for (int i = 1; i <= ix; i++) for (int j = 1; j <= il; j++)
#define F(x) ((double)(x##f))

/* ---------------------------------------------------------------- physical_constants.f90:31-37 */
double sigl[kx + 1], sigh[kx + 1], grdsig[kx + 1], grdscp[kx + 1];
FA<double, kx, 2> wvi;

/* ---------------------------------------------------------------- auxiliaries.f90:15-33 */
Grid2 precnv, precls, snowcv, snowls, cbmf, tsr, ssrd, ssr, slrd, slr, olr;
FA<double, ix, il, 3> slru, ustr, vstr, shf, evap, hfluxn;

/* ---------------------------------------------------------------- mod_radcon.f90:22-47 */
double albsea = F(0.07), albice = F(0.60), albsn = F(0.60), epslw = F(0.05), emisfc = F(0.98), ablco2_ref;
FA<double, 301, 4> fband;
#define FBAND(jt, jb) fband((jt)-99, (jb))
Grid2 alb_l, alb_s, albsfc, snowc;
FA<double, ix, il, kx, 4> tau2;
FA<double, ix, il, kx, 2> st4a;
FA<double, ix, il, 2> stratc;
FA<double, ix, il, 4> flux;

/* physics.f90:12-39 */
void initialize_physics() {
    sigh[0] = geo.hsg[1];
    for (int k = 1; k <= kx; k++) {
        sigl[k] = log(geo.fsg[k]);
        sigh[k] = geo.hsg[k + 1];
        grdsig[k] = grav / (geo.dhs[k] * p0);
        grdscp[k] = grdsig[k] / cp;
    }
    for (int k = 1; k <= kx - 1; k++) {
        wvi(k, 1) = 1. / (sigl[k + 1] - sigl[k]);
        wvi(k, 2) = (log(sigh[k]) - sigl[k]) * wvi(k, 1);
    }
    wvi(kx, 1) = 0.;
    wvi(kx, 2) = ((double)logf(0.99f) - sigl[kx]) * wvi(kx - 1, 1);   /* :38 log(0.99) is real32 */
}

/* ---------------------------------------------------------------- humidity.f90:44-78 */
void get_qsat(const double* ta, const double* ps_, double sig, double* qsat) {
    const double e0 = 6.108e-3;   /* _p literal: true double */
    const double c1 = F(17.269), c2 = F(21.875), t0 = F(273.16), t1 = F(35.86), t2 = F(7.66);
    const int N = ix * il;
    for (int q = 0; q < N; q++) {
        if (ta[q] >= t0) qsat[q] = e0 * exp(c1 * (ta[q] - t0) / (ta[q] - t1));
        else qsat[q] = e0 * exp(c2 * (ta[q] - t0) / (ta[q] - t2));
    }
    if (sig <= 0.0) {
        const double ps11 = ps_[0];
        for (int q = 0; q < N; q++) qsat[q] = 622.0 * qsat[q] / (ps11 - F(0.378) * qsat[q]);
    } else {
        for (int q = 0; q < N; q++) qsat[q] = 622.0 * qsat[q] / (sig * ps_[q] - F(0.378) * qsat[q]);
    }
}
/* humidity.f90:17-27 */
static void spec_hum_to_rel_hum(const double* ta, const double* ps_, double sig, const double* qa, double* rh, double* qsat) {
    get_qsat(ta, ps_, sig, qsat);
    for (int q = 0; q < ix * il; q++) rh[q] = qa[q] / qsat[q];
}

/* ---------------------------------------------------------------- convection.f90 */
static const double psmin = F(0.8), trcnv = 6.0, rhbl = F(0.9), rhil = F(0.7), entmax = 0.5, smf = F(0.8);

/* convection.f90:170-245 */
static void diagnose_convection(const Grid2& psa, const Grid3& se, const Grid3& qa, const Grid3& qsat, FA<int, ix, il>& itop, Grid2& qdif) {
    static Grid3 mss;   /* mss(ix,il,2:kx) */
    const int nl1 = kx - 1, nlp = kx + 1;
    for (int k = 2; k <= kx; k++)
        GLOOP mss(i, j, k) = se(i, j, k) + alhc * qsat(i, j, k);
    const double rlhc = 1.0 / alhc;
    IJLOOP {
        itop(i, j) = nlp;
        if (psa(i, j) > psmin) {
            double mse0 = se(i, j, kx) + alhc * qa(i, j, kx);
            double mse1 = se(i, j, nl1) + alhc * qa(i, j, nl1);
            mse1 = std::min(mse0, mse1);
            double mss0 = std::max(mse0, mss(i, j, kx));
            int ktop1 = kx, ktop2 = kx;
            double msthr = 0.0;
            for (int k = kx - 3; k >= 3; k--) {
                double mss2 = mss(i, j, k) + wvi(k, 2) * (mss(i, j, k + 1) - mss(i, j, k));
                if (mss0 > mss2) ktop1 = k;
                if (mse1 > mss2) { ktop2 = k; msthr = mss2; }
            }
            if (ktop1 < kx) {
                double qthr0 = rhbl * qsat(i, j, kx);
                double qthr1 = rhbl * qsat(i, j, nl1);
                bool lqthr = (qa(i, j, kx) > qthr0 && qa(i, j, nl1) > qthr1);
                if (ktop2 < kx) {
                    itop(i, j) = ktop1;
                    qdif(i, j) = std::max(qa(i, j, kx) - qthr0, (mse0 - msthr) * rlhc);
                } else if (lqthr) {
                    itop(i, j) = ktop1;
                    qdif(i, j) = qa(i, j, kx) - qthr0;
                }
            }
        }
    }
}

/* convection.f90:27-158 */
static void get_convection_tendencies(const Grid2& psa, const Grid3& se, const Grid3& qa, const Grid3& qsat, FA<int, ix, il>& itop,
                                      Grid2& cbmf_, Grid2& precnv_, Grid3& dfse, Grid3& dfqa) {
    static Grid2 qdif;
    double entr[kx + 1];
    const int nl1 = kx - 1, nlp = kx + 1;
    const double fqmax = 5.0;
    const double fm0 = p0 * geo.dhs[kx] / (grav * trcnv * 3600.0);
    const double rdps = 2.0 / (1.0 - psmin);
    dfse.fill(0.0); dfqa.fill(0.0); cbmf_.fill(0.0); precnv_.fill(0.0);
    double sentr = 0.0;
    for (int k = 2; k <= nl1; k++) {
        double e = std::max(0.0, geo.fsg[k] - 0.5);
        entr[k] = e * e;
        sentr = sentr + entr[k];
    }
    sentr = entmax / sentr;
    for (int k = 2; k <= nl1; k++) entr[k] = entr[k] * sentr;
    diagnose_convection(psa, se, qa, qsat, itop, qdif);
    IJLOOP {
        if (itop(i, j) == nlp) continue;
        int k = kx, k1 = k - 1;
        double qmax = std::max(F(1.01) * qa(i, j, k), qsat(i, j, k));
        double sb = se(i, j, k1) + wvi(k1, 2) * (se(i, j, k) - se(i, j, k1));
        double qb = qa(i, j, k1) + wvi(k1, 2) * (qa(i, j, k) - qa(i, j, k1));
        qb = std::min(qb, qa(i, j, k));
        double fpsa = psa(i, j) * std::min(1.0, (psa(i, j) - psmin) * rdps);
        double fmass = fm0 * fpsa * std::min(fqmax, qdif(i, j) / (qmax - qb));
        cbmf_(i, j) = fmass;
        double fus = fmass * se(i, j, k);
        double fuq = fmass * qmax;
        double fds = fmass * sb;
        double fdq = fmass * qb;
        dfse(i, j, k) = fds - fus;
        dfqa(i, j, k) = fdq - fuq;
        for (k = kx - 1; k >= itop(i, j) + 1; k--) {
            k1 = k - 1;
            dfse(i, j, k) = fus - fds;
            dfqa(i, j, k) = fuq - fdq;
            double enmass = entr[k] * psa(i, j) * cbmf_(i, j);
            fmass = fmass + enmass;
            fus = fus + enmass * se(i, j, k);
            fuq = fuq + enmass * qa(i, j, k);
            sb = se(i, j, k1) + wvi(k1, 2) * (se(i, j, k) - se(i, j, k1));
            qb = qa(i, j, k1) + wvi(k1, 2) * (qa(i, j, k) - qa(i, j, k1));
            fds = fmass * sb;
            fdq = fmass * qb;
            dfse(i, j, k) = dfse(i, j, k) + fds - fus;
            dfqa(i, j, k) = dfqa(i, j, k) + fdq - fuq;
            double delq = rhil * qsat(i, j, k) - qa(i, j, k);
            if (delq > 0.0) {
                double fsq = smf * cbmf_(i, j) * delq;
                dfqa(i, j, k) = dfqa(i, j, k) + fsq;
                dfqa(i, j, kx) = dfqa(i, j, kx) - fsq;
            }
        }
        k = itop(i, j);
        double qsatb = qsat(i, j, k) + wvi(k, 2) * (qsat(i, j, k + 1) - qsat(i, j, k));
        precnv_(i, j) = std::max(fuq - fmass * qsatb, 0.0);
        dfse(i, j, k) = fus - fds + alhc * precnv_(i, j);
        dfqa(i, j, k) = fuq - fdq - precnv_(i, j);
    }
}

/* ---------------------------------------------------------------- large_scale_condensation.f90:33-95 */
static void get_large_scale_condensation_tendencies(const Grid2& psa, const Grid3& qa, const Grid3& qsat, FA<int, ix, il>& itop,
                                                    Grid2& precls_, Grid3& dtlsc, Grid3& dqlsc) {
    const double trlsc = 4.0, rhlsc = F(0.9), drhlsc = F(0.1), rhblsc = F(0.95);
    static Grid2 psa2;
    const double qsmax = 10.0;
    const double rtlsc = 1.0 / (trlsc * 3600.0);
    const double tfact = alhc / cp;
    const double prg = p0 / grav;
    GLOOP { dtlsc(i, j, 1) = 0.0; dqlsc(i, j, 1) = 0.0; }
    precls_.fill(0.0);
    GLOOP psa2(i, j) = psa(i, j) * psa(i, j);
    for (int k = 2; k <= kx; k++) {
        double sig2 = geo.fsg[k] * geo.fsg[k];
        double rhref = rhlsc + drhlsc * (sig2 - 1.0);
        if (k == kx) rhref = std::max(rhref, rhblsc);
        double dqmax = qsmax * sig2 * rtlsc;
        IJLOOP {
            double dqa = rhref * qsat(i, j, k) - qa(i, j, k);
            if (dqa < 0.0) {
                itop(i, j) = std::min(k, itop(i, j));
                dqlsc(i, j, k) = dqa * rtlsc;
                dtlsc(i, j, k) = tfact * std::min(-dqlsc(i, j, k), dqmax * psa2(i, j));
            } else {
                dqlsc(i, j, k) = 0.0;
                dtlsc(i, j, k) = 0.0;
            }
        }
    }
    for (int k = 2; k <= kx; k++) {
        double pfact = geo.dhs[k] * prg;
        GLOOP precls_(i, j) = precls_(i, j) - pfact * dqlsc(i, j, k);
    }
    GLOOP precls_(i, j) = precls_(i, j) * psa(i, j);
}

/* ---------------------------------------------------------------- shortwave_radiation.f90 */
static const double solc = 342.0, rhcl1 = F(0.30), rhcl2 = 1.00, qacl = F(0.20), wpcl = F(0.2), pmaxcl = 10.0;
static const double clsmax = F(0.60), clsminl = F(0.15), gse_s0 = 0.25, gse_s1 = F(0.40);
static const double albcl = F(0.43), albcls = 0.50, epssw = F(0.020);
static const double absdry = F(0.033), absaer = F(0.033), abswv1 = F(0.022), abswv2 = 15.000, abscl1 = F(0.015), abscl2 = F(0.15);
static const double ablwin = F(0.3), ablwv1 = F(0.7), ablwv2 = 50.0, ablcl1 = 12.0, ablcl2 = F(0.6);
double ablco2 = 6.0;
Grid2 fsol, ozone, ozupp, zenit, stratz, qcloud;
bool compute_shortwave = true;

/* shortwave_radiation.f90:74-234; icltop is the first (ix,il) slice of physics' icltop(ix,il,2) */
static void get_shortwave_rad_fluxes(const Grid2& psa, const Grid3& qa, const FA<int, ix, il, 2>& icltop, const Grid2& cloudc, const Grid2& clstr,
                                     Grid2& fsfcd, Grid2& fsfc, Grid2& ftop, Grid3& dfabs) {
    static Grid2 acloud, psaz;
    const int nl1 = kx - 1;
    const double fband2 = F(0.05);
    const double fband1 = 1.0 - fband2;
    tau2.fill(0.0);
    IJLOOP {
        if (icltop(i, j, 1) <= kx) tau2(i, j, icltop(i, j, 1), 3) = albcl * cloudc(i, j);
        tau2(i, j, kx, 3) = albcls * clstr(i, j);
    }
    GLOOP psaz(i, j) = psa(i, j) * zenit(i, j);
    GLOOP acloud(i, j) = cloudc(i, j) * std::min(abscl1 * qcloud(i, j), abscl2);
    GLOOP tau2(i, j, 1, 1) = exp(-psaz(i, j) * geo.dhs[1] * absdry);
    for (int k = 2; k <= nl1; k++) {
        double abs1 = absdry + absaer * (geo.fsg[k] * geo.fsg[k]);
        IJLOOP {
            if (k >= icltop(i, j, 1)) tau2(i, j, k, 1) = exp(-psaz(i, j) * geo.dhs[k] * (abs1 + abswv1 * qa(i, j, k) + acloud(i, j)));
            else tau2(i, j, k, 1) = exp(-psaz(i, j) * geo.dhs[k] * (abs1 + abswv1 * qa(i, j, k)));
        }
    }
    {
        double abs1 = absdry + absaer * (geo.fsg[kx] * geo.fsg[kx]);
        GLOOP tau2(i, j, kx, 1) = exp(-psaz(i, j) * geo.dhs[kx] * (abs1 + abswv1 * qa(i, j, kx)));
    }
    for (int k = 2; k <= kx; k++)
        GLOOP tau2(i, j, k, 2) = exp(-psaz(i, j) * geo.dhs[k] * abswv2 * qa(i, j, k));
    /* 3. downward flux */
    GLOOP {
        ftop(i, j) = fsol(i, j);
        flux(i, j, 1) = fsol(i, j) * fband1;
        flux(i, j, 2) = fsol(i, j) * fband2;
    }
    GLOOP {
        int k = 1;
        dfabs(i, j, k) = flux(i, j, 1);
        flux(i, j, 1) = tau2(i, j, k, 1) * (flux(i, j, 1) - ozupp(i, j) * psa(i, j));
        dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, 1);
        k = 2;
        dfabs(i, j, k) = flux(i, j, 1);
        flux(i, j, 1) = tau2(i, j, k, 1) * (flux(i, j, 1) - ozone(i, j) * psa(i, j));
        dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, 1);
    }
    for (int k = 3; k <= kx; k++)
        GLOOP {
            tau2(i, j, k, 3) = flux(i, j, 1) * tau2(i, j, k, 3);
            flux(i, j, 1) = flux(i, j, 1) - tau2(i, j, k, 3);
            dfabs(i, j, k) = flux(i, j, 1);
            flux(i, j, 1) = tau2(i, j, k, 1) * flux(i, j, 1);
            dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, 1);
        }
    for (int k = 2; k <= kx; k++)
        GLOOP {
            dfabs(i, j, k) = dfabs(i, j, k) + flux(i, j, 2);
            flux(i, j, 2) = tau2(i, j, k, 2) * flux(i, j, 2);
            dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, 2);
        }
    /* 4. upward flux */
    GLOOP {
        fsfcd(i, j) = flux(i, j, 1) + flux(i, j, 2);
        flux(i, j, 1) = flux(i, j, 1) * albsfc(i, j);
        fsfc(i, j) = fsfcd(i, j) - flux(i, j, 1);
    }
    for (int k = kx; k >= 1; k--)
        GLOOP {
            dfabs(i, j, k) = dfabs(i, j, k) + flux(i, j, 1);
            flux(i, j, 1) = tau2(i, j, k, 1) * flux(i, j, 1);
            dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, 1);
            flux(i, j, 1) = flux(i, j, 1) + tau2(i, j, k, 3);
        }
    GLOOP ftop(i, j) = ftop(i, j) - flux(i, j, 1);
    /* 5. longwave transmissivities */
    {
        int k = 1;
        GLOOP {
            tau2(i, j, k, 1) = exp(-psa(i, j) * geo.dhs[k] * ablwin);
            tau2(i, j, k, 2) = exp(-psa(i, j) * geo.dhs[k] * ablco2);
            tau2(i, j, k, 3) = 1.0;
            tau2(i, j, k, 4) = 1.0;
        }
    }
    for (int k = 2; k <= kx; k += kx - 2)
        GLOOP {
            tau2(i, j, k, 1) = exp(-psa(i, j) * geo.dhs[k] * ablwin);
            tau2(i, j, k, 2) = exp(-psa(i, j) * geo.dhs[k] * ablco2);
            tau2(i, j, k, 3) = exp(-psa(i, j) * geo.dhs[k] * ablwv1 * qa(i, j, k));
            tau2(i, j, k, 4) = exp(-psa(i, j) * geo.dhs[k] * ablwv2 * qa(i, j, k));
        }
    GLOOP acloud(i, j) = cloudc(i, j) * ablcl2;
    for (int k = 3; k <= nl1; k++)
        IJLOOP {
            double deltap = psa(i, j) * geo.dhs[k];
            double acloud1;
            if (k < icltop(i, j, 1)) acloud1 = acloud(i, j);
            else acloud1 = ablcl1 * cloudc(i, j);
            tau2(i, j, k, 1) = exp(-deltap * (ablwin + acloud1));
            tau2(i, j, k, 2) = exp(-deltap * ablco2);
            tau2(i, j, k, 3) = exp(-deltap * std::max(ablwv1 * qa(i, j, k), acloud(i, j)));
            tau2(i, j, k, 4) = exp(-deltap * std::max(ablwv2 * qa(i, j, k), acloud(i, j)));
        }
    double eps1 = epslw / (geo.dhs[1] + geo.dhs[2]);
    GLOOP {
        stratc(i, j, 1) = stratz(i, j) * psa(i, j);
        stratc(i, j, 2) = eps1 * psa(i, j);
    }
}

/* shortwave_radiation.f90:287-329 */
static void solar(double tyear_, double csol, double* topsr /*1-based*/) {
    const double pigr = (double)(2.0f * asinf(1.0f));   /* :300 real32 */
    double alpha = 2.0 * pigr * tyear_;
    double ca1 = cos(alpha), sa1 = sin(alpha);
    double ca2 = ca1 * ca1 - sa1 * sa1;
    double sa2 = 2. * sa1 * ca1;
    double ca3 = ca1 * ca2 - sa1 * sa2;
    double sa3 = sa1 * ca2 + sa2 * ca1;
    double decl = F(0.006918) - F(0.399912) * ca1 + F(0.070257) * sa1 - F(0.006758) * ca2 + F(0.000907) * sa2
                  - F(0.002697) * ca3 + F(0.001480) * sa3;
    double fdis = F(1.000110) + F(0.034221) * ca1 + F(0.001280) * sa1 + F(0.000719) * ca2 + F(0.000077) * sa2;
    double cdecl_ = cos(decl), sdecl = sin(decl);
    double tdecl = sdecl / cdecl_;
    double csolp = csol / pigr;
    for (int j = 1; j <= il; j++) {
        double ch0 = std::min(1.0, std::max(-1.0, -tdecl * geo.sia[j] / geo.coa[j]));
        double h0 = acos(ch0);
        double sh0 = sin(h0);
        topsr[j] = csolp * fdis * (h0 * geo.sia[j] * sdecl + sh0 * geo.coa[j] * cdecl_);
    }
}

/* shortwave_radiation.f90:238-284 */
void get_zonal_average_fields(double tyear_) {
    double topsr[il + 1];
    /* :248 4.0*asin(1.0) and 10.0/365.0 are real32 */
    double alpha = (double)(4.0f * asinf(1.0f)) * (tyear_ + (double)(10.0f / 365.0f));
    double dalpha = 0.0;
    double coz1 = 1.0 * std::max(0.0, cos(alpha - dalpha));
    double coz2 = F(1.8);
    double azen = 1.0;
    /* :257 -cos(alpha)*23.45*asin(1.0)/90.0, left to right with real32 asin */
    double rzen = -cos(alpha) * F(23.45) * (double)asinf(1.0f) / 90.0;
    double fs0 = 6.0;
    solar(tyear_, 4.0 * solc, topsr);
    for (int j = 1; j <= il; j++) {
        double flat2 = 1.5 * (geo.sia[j] * geo.sia[j]) - 0.5;
        double z = 1.0 - (geo.coa[j] * cos(rzen) + geo.sia[j] * sin(rzen));
        for (int i = 1; i <= ix; i++) {
            fsol(i, j) = topsr[j];
            ozupp(i, j) = 0.5 * epssw;
            ozone(i, j) = F(0.4) * epssw * (1.0 + coz1 * geo.sia[j] + coz2 * flat2);
            zenit(i, j) = 1.0 + azen * (z * z);   /* (...)**nzen, nzen = 2 real */
            ozupp(i, j) = fsol(i, j) * ozupp(i, j) * zenit(i, j);
            ozone(i, j) = fsol(i, j) * ozone(i, j) * zenit(i, j);
            stratz(i, j) = std::max(fs0 - fsol(i, j), 0.0);
        }
    }
}

/* shortwave_radiation.f90:332-410 */
static void clouds(const Grid3& qa, const Grid3& rh, const Grid2& precnv_, const Grid2& precls_, const FA<int, ix, il>& iptop, const Grid2& gse,
                   const Grid2& fmask_, FA<int, ix, il, 2>& icltop, Grid2& cloudc, Grid2& clstr) {
    const int nl1 = kx - 1, nlp = kx + 1;
    const double rrcl = 1. / (rhcl2 - rhcl1);
    IJLOOP {
        if (rh(i, j, nl1) > rhcl1) {
            cloudc(i, j) = rh(i, j, nl1) - rhcl1;
            icltop(i, j, 1) = nl1;
        } else {
            cloudc(i, j) = 0.0;
            icltop(i, j, 1) = nlp;
        }
    }
    for (int k = 3; k <= kx - 2; k++)
        IJLOOP {
            double drh = rh(i, j, k) - rhcl1;
            if (drh > cloudc(i, j) && qa(i, j, k) > qacl) {
                cloudc(i, j) = drh;
                icltop(i, j, 1) = k;
            }
        }
    IJLOOP {
        double pr1 = std::min(pmaxcl, F(86.4) * (precnv_(i, j) + precls_(i, j)));
        double c = std::min(1.0, cloudc(i, j) * rrcl);
        cloudc(i, j) = std::min(1.0, wpcl * sqrt(pr1) + c * c);
        icltop(i, j, 1) = std::min(iptop(i, j), icltop(i, j, 1));
    }
    GLOOP qcloud(i, j) = qa(i, j, nl1);
    const double clfact = F(1.2);
    const double rgse = 1.0 / (gse_s1 - gse_s0);
    IJLOOP {
        double fstab = std::max(0.0, std::min(1.0, rgse * (gse(i, j) - gse_s0)));
        clstr(i, j) = fstab * std::max(clsmax - clfact * cloudc(i, j), 0.0);
        double clstrl = std::max(clstr(i, j), clsminl) * rh(i, j, kx);
        clstr(i, j) = clstr(i, j) + fmask_(i, j) * (clstrl - clstr(i, j));
    }
}

/* ---------------------------------------------------------------- longwave_radiation.f90 */
static inline int nint_(double x) { return (int)lround(x); }

/* longwave_radiation.f90:197-220 */
void radset() {
    double eps1 = 1.0 - epslw;
    for (int jtemp = 200; jtemp <= 320; jtemp++) {
        /* brackets are real32 (real32 literal * integer) */
        float d2 = (float)((jtemp - 247) * (jtemp - 247)), d3 = (float)((jtemp - 282) * (jtemp - 282)), d4 = (float)((jtemp - 315) * (jtemp - 315));
        FBAND(jtemp, 2) = (double)(0.148f - 3.0e-6f * d2) * eps1;
        FBAND(jtemp, 3) = (double)(0.356f - 5.2e-6f * d3) * eps1;
        FBAND(jtemp, 4) = (double)(0.314f + 1.0e-5f * d4) * eps1;
        FBAND(jtemp, 1) = eps1 - (FBAND(jtemp, 2) + FBAND(jtemp, 3) + FBAND(jtemp, 4));
    }
    for (int jb = 1; jb <= 4; jb++) {
        for (int jtemp = 100; jtemp <= 199; jtemp++) FBAND(jtemp, jb) = FBAND(200, jb);
        for (int jtemp = 321; jtemp <= 400; jtemp++) FBAND(jtemp, jb) = FBAND(320, jb);
    }
}

/* longwave_radiation.f90:16-117 */
static void get_downward_longwave_rad_fluxes(const Grid3& ta, Grid2& fsfcd, Grid3& dfabs) {
    const int nl1 = kx - 1, nband = 4;
    for (int k = 1; k <= nl1; k++)
        GLOOP st4a(i, j, k, 1) = ta(i, j, k) + wvi(k, 2) * (ta(i, j, k + 1) - ta(i, j, k));
    GLOOP {
        st4a(i, j, 1, 2) = 0.75 * ta(i, j, 1) + 0.25 * st4a(i, j, 1, 1);
        st4a(i, j, 2, 2) = 0.50 * ta(i, j, 2) + 0.25 * (st4a(i, j, 1, 1) + st4a(i, j, 2, 1));
    }
    const double anis = 1.0;
    for (int k = 3; k <= nl1; k++)
        GLOOP st4a(i, j, k, 2) = 0.5 * anis * std::max(st4a(i, j, k, 1) - st4a(i, j, k - 1, 1), 0.0);
    GLOOP st4a(i, j, kx, 2) = anis * std::max(ta(i, j, kx) - st4a(i, j, nl1, 1), 0.0);
    for (int k = 1; k <= 2; k++)
        GLOOP {
            double x = st4a(i, j, k, 2);
            st4a(i, j, k, 1) = sbc * ((x * x) * (x * x));
            st4a(i, j, k, 2) = 0.0;
        }
    for (int k = 3; k <= kx; k++)
        GLOOP {
            double x = ta(i, j, k);
            double st3a = sbc * ((x * x) * x);
            st4a(i, j, k, 1) = st3a * ta(i, j, k);
            st4a(i, j, k, 2) = 4.0 * st3a * st4a(i, j, k, 2);
        }
    fsfcd.fill(0.0);
    dfabs.fill(0.0);
    {
        int k = 1;
        for (int jb = 1; jb <= 2; jb++)
            IJLOOP {
                double emis = 1.0 - tau2(i, j, k, jb);
                double brad = FBAND(nint_(ta(i, j, k)), jb) * (st4a(i, j, k, 1) + emis * st4a(i, j, k, 2));
                flux(i, j, jb) = emis * brad;
                dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, jb);
            }
    }
    for (int jb = 3; jb <= nband; jb++)
        GLOOP flux(i, j, jb) = 0.0;
    for (int jb = 1; jb <= nband; jb++)
        for (int k = 2; k <= kx; k++)
            IJLOOP {
                double emis = 1.0 - tau2(i, j, k, jb);
                double brad = FBAND(nint_(ta(i, j, k)), jb) * (st4a(i, j, k, 1) + emis * st4a(i, j, k, 2));
                dfabs(i, j, k) = dfabs(i, j, k) + flux(i, j, jb);
                flux(i, j, jb) = tau2(i, j, k, jb) * flux(i, j, jb) + emis * brad;
                dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, jb);
            }
    for (int jb = 1; jb <= nband; jb++)
        GLOOP fsfcd(i, j) = fsfcd(i, j) + emisfc * flux(i, j, jb);
    GLOOP {
        double corlw = epslw * emisfc * st4a(i, j, kx, 1);
        dfabs(i, j, kx) = dfabs(i, j, kx) - corlw;
        fsfcd(i, j) = fsfcd(i, j) + corlw;
    }
}

/* longwave_radiation.f90:120-194; fsfcu is slru(:,:,3) */
static void get_upward_longwave_rad_fluxes(const Grid3& ta, const Grid2& ts, const Grid2& fsfcd, const double* fsfcu, Grid2& fsfc, Grid2& ftop, Grid3& dfabs) {
    const int nband = 4;
#define FSFCU(i, j) fsfcu[((i)-1) + (size_t)ix * ((j)-1)]
    const double refsfc = 1.0 - emisfc;
    GLOOP fsfc(i, j) = FSFCU(i, j) - fsfcd(i, j);
    for (int jb = 1; jb <= nband; jb++)
        IJLOOP flux(i, j, jb) = FBAND(nint_(ts(i, j)), jb) * FSFCU(i, j) + refsfc * flux(i, j, jb);
    GLOOP dfabs(i, j, kx) = dfabs(i, j, kx) + epslw * FSFCU(i, j);
    for (int jb = 1; jb <= nband; jb++)
        for (int k = kx; k >= 2; k--)
            IJLOOP {
                double emis = 1.0 - tau2(i, j, k, jb);
                double brad = FBAND(nint_(ta(i, j, k)), jb) * (st4a(i, j, k, 1) - emis * st4a(i, j, k, 2));
                dfabs(i, j, k) = dfabs(i, j, k) + flux(i, j, jb);
                flux(i, j, jb) = tau2(i, j, k, jb) * flux(i, j, jb) + emis * brad;
                dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, jb);
            }
    {
        int k = 1;
        for (int jb = 1; jb <= 2; jb++)
            IJLOOP {
                double emis = 1.0 - tau2(i, j, k, jb);
                double brad = FBAND(nint_(ta(i, j, k)), jb) * (st4a(i, j, k, 1) - emis * st4a(i, j, k, 2));
                dfabs(i, j, k) = dfabs(i, j, k) + flux(i, j, jb);
                flux(i, j, jb) = tau2(i, j, k, jb) * flux(i, j, jb) + emis * brad;
                dfabs(i, j, k) = dfabs(i, j, k) - flux(i, j, jb);
            }
    }
    GLOOP {
        double corlw1 = geo.dhs[1] * stratc(i, j, 2) * st4a(i, j, 1, 1) + stratc(i, j, 1);
        double corlw2 = geo.dhs[2] * stratc(i, j, 2) * st4a(i, j, 2, 1);
        dfabs(i, j, 1) = dfabs(i, j, 1) - corlw1;
        dfabs(i, j, 2) = dfabs(i, j, 2) - corlw2;
        ftop(i, j) = corlw1 + corlw2;
    }
    for (int jb = 1; jb <= nband; jb++)
        GLOOP ftop(i, j) = ftop(i, j) + flux(i, j, jb);
#undef FSFCU
}

/* ---------------------------------------------------------------- surface_fluxes.f90 */
static const double fwind0 = F(0.95), ftemp0 = 1.0, fhum0 = 0.0, cdl = F(2.4e-3), cds = F(1.0e-3), chl = F(1.2e-3), chs = F(0.9e-3);
static const double vgust = 5.0, ctday = F(1.0e-2), dtheta = 3.0, fstab_ = F(0.67), hdrag = 2000.0, clambda = 7.0, clambsn = 7.0;
Grid2 forog;

/* surface_fluxes.f90:300-309 */
void set_orog_land_sfc_drag(const Grid2& phi0_) {
    double rhdrag = 1.0 / (grav * hdrag);
    GLOOP forog(i, j) = 1.0 + rhdrag * (1.0 - exp(-std::max(phi0_(i, j), 0.0) * rhdrag));
}

/* surface_fluxes.f90:42-296 (lfluxland = .true.; the sea_coupling_flag > 0 re-call is not reachable: sea_model.f90:185-190 stops) */
static void get_surface_fluxes(const Grid2& psa, const Grid3& ua, const Grid3& va, const Grid3& ta, const Grid3& qa, const Grid3& rh, const Grid3& phi_,
                               const Grid2& phi0_, const Grid2& fmask_, const Grid2& tsea, const Grid2& ssrd_, const Grid2& slrd_,
                               FA<double, ix, il, 3>& ustr_, FA<double, ix, il, 3>& vstr_, FA<double, ix, il, 3>& shf_, FA<double, ix, il, 3>& evap_,
                               FA<double, ix, il, 3>& slru_, FA<double, ix, il, 3>& hfluxn_, Grid2& tsfc, Grid2& tskin, Grid2& u0, Grid2& v0, Grid2& t0,
                               bool lfluxland) {
    static FA<double, ix, il, 2> t1, q1, t2, qsat0;
    static FA<double, ix, il, 3> denvvs;   /* denvvs(ix,il,0:2) -> third index +1 */
    static Grid2 dslr, dtskin, clamb, cdsdv, tsk3;
    const bool lscasym = true, lskineb = true;
    const double esbc = emisfc * sbc;
    const double ghum0 = 1.0 - fhum0;
    int ks = 2;
    if (lfluxland) {
        GLOOP { u0(i, j) = fwind0 * ua(i, j, kx); v0(i, j) = fwind0 * va(i, j, kx); }
        const double gtemp0 = 1.0 - ftemp0;
        const double rcp = 1.0 / cp;
        const int nl1 = kx - 1;
        IJLOOP {
            double dt1 = wvi(kx, 2) * (ta(i, j, kx) - ta(i, j, nl1));
            t1(i, j, 1) = ta(i, j, kx) + dt1;
            t1(i, j, 2) = t1(i, j, 1) - phi0_(i, j) * dt1 / (rgas * 288.0 * sigl[kx]);
            t2(i, j, 2) = ta(i, j, kx) + rcp * phi_(i, j, kx);
            t2(i, j, 1) = t2(i, j, 2) - rcp * phi0_(i, j);
        }
        IJLOOP {
            if (ta(i, j, kx) > ta(i, j, nl1)) {
                t1(i, j, 1) = ftemp0 * t1(i, j, 1) + gtemp0 * t2(i, j, 1);
                t1(i, j, 2) = ftemp0 * t1(i, j, 2) + gtemp0 * t2(i, j, 2);
            } else {
                t1(i, j, 1) = ta(i, j, kx);
                t1(i, j, 2) = ta(i, j, kx);
            }
            t0(i, j) = t1(i, j, 2) + fmask_(i, j) * (t1(i, j, 1) - t1(i, j, 2));
        }
        GLOOP denvvs(i, j, 1) = (p0 * psa(i, j) / (rgas * t0(i, j))) * sqrt(u0(i, j) * u0(i, j) + v0(i, j) * v0(i, j) + vgust * vgust);
        for (int j = 1; j <= il; j++)
            for (int i = 1; i <= ix; i++)
                tskin(i, j) = stl_am(i, j) + ctday * sqrt(geo.coa[j]) * ssrd_(i, j) * (1.0 - alb_l(i, j)) * psa(i, j);
        double rdth = fstab_ / dtheta;
        double astab = 1.0;
        if (lscasym) astab = 0.5;
        IJLOOP {
            double dthl;
            if (tskin(i, j) > t2(i, j, 1)) dthl = std::min(dtheta, tskin(i, j) - t2(i, j, 1));
            else dthl = std::max(-dtheta, astab * (tskin(i, j) - t2(i, j, 1)));
            denvvs(i, j, 2) = denvvs(i, j, 1) * (1.0 + dthl * rdth);
        }
        IJLOOP {
            double cdldv = cdl * denvvs(i, j, 1) * forog(i, j);
            ustr_(i, j, 1) = -cdldv * ua(i, j, kx);
            vstr_(i, j, 1) = -cdldv * va(i, j, kx);
        }
        const double chlcp = chl * cp;
        GLOOP shf_(i, j, 1) = chlcp * denvvs(i, j, 2) * (tskin(i, j) - t1(i, j, 1));
        GLOOP q1(i, j, 1) = qa(i, j, kx);   /* fhum0 = 0 */
        get_qsat(tskin.p(), psa.p(), 1.0, qsat0.p(1, 1, 1));
        GLOOP evap_(i, j, 1) = chl * denvvs(i, j, 2) * std::max(0.0, soilw_am(i, j) * qsat0(i, j, 1) - q1(i, j, 1));
        GLOOP {
            tsk3(i, j) = (tskin(i, j) * tskin(i, j)) * tskin(i, j);
            dslr(i, j) = 4.0 * esbc * tsk3(i, j);
            slru_(i, j, 1) = esbc * tsk3(i, j) * tskin(i, j);
            hfluxn_(i, j, 1) = ssrd_(i, j) * (1.0 - alb_l(i, j)) + slrd_(i, j) - (slru_(i, j, 1) + shf_(i, j, 1) + alhc * evap_(i, j, 1));
        }
        if (lskineb) {
            GLOOP {
                clamb(i, j) = clambda + snowc(i, j) * (clambsn - clambda);
                hfluxn_(i, j, 1) = hfluxn_(i, j, 1) - clamb(i, j) * (tskin(i, j) - stl_am(i, j));
                dtskin(i, j) = tskin(i, j) + 1.0;
            }
            get_qsat(dtskin.p(), psa.p(), 1.0, qsat0.p(1, 1, 2));
            IJLOOP {
                if (evap_(i, j, 1) > 0.0) qsat0(i, j, 2) = soilw_am(i, j) * (qsat0(i, j, 2) - qsat0(i, j, 1));
                else qsat0(i, j, 2) = 0.0;
            }
            GLOOP {
                dtskin(i, j) = hfluxn_(i, j, 1) / (clamb(i, j) + dslr(i, j) + chl * denvvs(i, j, 2) * (cp + alhc * qsat0(i, j, 2)));
                tskin(i, j) = tskin(i, j) + dtskin(i, j);
                shf_(i, j, 1) = shf_(i, j, 1) + chlcp * denvvs(i, j, 2) * dtskin(i, j);
                evap_(i, j, 1) = evap_(i, j, 1) + chl * denvvs(i, j, 2) * qsat0(i, j, 2) * dtskin(i, j);
                slru_(i, j, 1) = slru_(i, j, 1) + dslr(i, j) * dtskin(i, j);
                hfluxn_(i, j, 1) = clamb(i, j) * (tskin(i, j) - stl_am(i, j));
            }
        }
        rdth = fstab_ / dtheta;
        astab = 1.0;
        if (lscasym) astab = 0.5;
        IJLOOP {
            double dths;
            if (tsea(i, j) > t2(i, j, 2)) dths = std::min(dtheta, tsea(i, j) - t2(i, j, 2));
            else dths = std::max(-dtheta, astab * (tsea(i, j) - t2(i, j, 2)));
            denvvs(i, j, 3) = denvvs(i, j, 1) * (1.0 + dths * rdth);
        }
        GLOOP q1(i, j, 2) = qa(i, j, kx);
        ks = 2;
        GLOOP {
            cdsdv(i, j) = cds * denvvs(i, j, ks + 1);
            ustr_(i, j, 2) = -cdsdv(i, j) * ua(i, j, kx);
            vstr_(i, j, 2) = -cdsdv(i, j) * va(i, j, kx);
        }
    }
    /* sea surface */
    GLOOP shf_(i, j, 2) = chs * cp * denvvs(i, j, ks + 1) * (tsea(i, j) - t1(i, j, 2));
    get_qsat(tsea.p(), psa.p(), 1.0, qsat0.p(1, 1, 2));
    GLOOP evap_(i, j, 2) = chs * denvvs(i, j, ks + 1) * (qsat0(i, j, 2) - q1(i, j, 2));
    GLOOP {
        double x = tsea(i, j);
        slru_(i, j, 2) = esbc * ((x * x) * (x * x));
        hfluxn_(i, j, 2) = ssrd_(i, j) * (1.0 - alb_s(i, j)) + slrd_(i, j) - slru_(i, j, 2) + shf_(i, j, 2) + alhc * evap_(i, j, 2);
    }
    if (lfluxland) {
        GLOOP {
            ustr_(i, j, 3) = ustr_(i, j, 2) + fmask_(i, j) * (ustr_(i, j, 1) - ustr_(i, j, 2));
            vstr_(i, j, 3) = vstr_(i, j, 2) + fmask_(i, j) * (vstr_(i, j, 1) - vstr_(i, j, 2));
            shf_(i, j, 3) = shf_(i, j, 2) + fmask_(i, j) * (shf_(i, j, 1) - shf_(i, j, 2));
            evap_(i, j, 3) = evap_(i, j, 2) + fmask_(i, j) * (evap_(i, j, 1) - evap_(i, j, 2));
            slru_(i, j, 3) = slru_(i, j, 2) + fmask_(i, j) * (slru_(i, j, 1) - slru_(i, j, 2));
            tsfc(i, j) = tsea(i, j) + fmask_(i, j) * (stl_am(i, j) - tsea(i, j));
            tskin(i, j) = tsea(i, j) + fmask_(i, j) * (tskin(i, j) - tsea(i, j));
            t0(i, j) = t1(i, j, 2) + fmask_(i, j) * (t1(i, j, 1) - t1(i, j, 2));
        }
    }
}

/* ---------------------------------------------------------------- vertical_diffusion.f90:30-143 */
static void get_vertical_diffusion_tend(const Grid3& se, const Grid3& rh, const Grid3& qa, const Grid3& qsat, const Grid3& phi_, const FA<int, ix, il>& icnv,
                                        Grid3& utenvd, Grid3& vtenvd, Grid3& ttenvd, Grid3& qtenvd) {
    const double trshc = 6.0, trvdi = 24.0, trvds = 6.0, redshc = 0.5, rhgrad = 0.5, segrad = F(0.1);
    double rsig[kx + 1], rsig1[kx + 1];
    const int nl1 = kx - 1;
    const double cshc = geo.dhs[kx] / 3600.0;
    const double cvdi = (sigh[nl1] - sigh[1]) / ((nl1 - 1) * 3600.0);
    const double fshcq = cshc / trshc;
    const double fshcse = cshc / (trshc * cp);
    const double fvdiq = cvdi / trvdi;
    const double fvdise = cvdi / (trvds * cp);
    for (int k = 1; k <= nl1; k++) {
        rsig[k] = 1.0 / geo.dhs[k];
        rsig1[k] = 1.0 / (1.0 - sigh[k]);
    }
    rsig[kx] = 1.0 / geo.dhs[kx];
    utenvd.fill(0.0); vtenvd.fill(0.0); ttenvd.fill(0.0); qtenvd.fill(0.0);
    double drh0 = rhgrad * (geo.fsg[kx] - geo.fsg[nl1]);
    double fvdiq2 = fvdiq * sigh[nl1];
    IJLOOP {
        double dmse = se(i, j, kx) - se(i, j, nl1) + alhc * (qa(i, j, kx) - qsat(i, j, nl1));
        double drh = rh(i, j, kx) - rh(i, j, nl1);
        double fcnv = 1.0;
        if (dmse >= 0.0) {
            if (icnv(i, j) > 0) fcnv = redshc;
            double fluxse = fcnv * fshcse * dmse;
            ttenvd(i, j, nl1) = fluxse * rsig[nl1];
            ttenvd(i, j, kx) = -fluxse * rsig[kx];
            if (drh >= 0.0) {
                double fluxq = fcnv * fshcq * qsat(i, j, kx) * drh;
                qtenvd(i, j, nl1) = fluxq * rsig[nl1];
                qtenvd(i, j, kx) = -fluxq * rsig[kx];
            }
        } else if (drh > drh0) {
            double fluxq = fvdiq2 * qsat(i, j, nl1) * drh;
            qtenvd(i, j, nl1) = fluxq * rsig[nl1];
            qtenvd(i, j, kx) = -fluxq * rsig[kx];
        }
    }
    for (int k = 3; k <= kx - 2; k++) {
        if (sigh[k] > 0.5) {
            drh0 = rhgrad * (geo.fsg[k + 1] - geo.fsg[k]);
            fvdiq2 = fvdiq * sigh[k];
            IJLOOP {
                double drh = rh(i, j, k + 1) - rh(i, j, k);
                if (drh >= drh0) {
                    double fluxq = fvdiq2 * qsat(i, j, k) * drh;
                    qtenvd(i, j, k) = qtenvd(i, j, k) + fluxq * rsig[k];
                    qtenvd(i, j, k + 1) = qtenvd(i, j, k + 1) - fluxq * rsig[k + 1];
                }
            }
        }
    }
    for (int k = 1; k <= nl1; k++)
        IJLOOP {
            double se0 = se(i, j, k + 1) + segrad * (phi_(i, j, k) - phi_(i, j, k + 1));
            if (se(i, j, k) < se0) {
                double fluxse = fvdise * (se0 - se(i, j, k));
                ttenvd(i, j, k) = ttenvd(i, j, k) + fluxse * rsig[k];
                for (int k1 = k + 1; k1 <= kx; k1++) ttenvd(i, j, k1) = ttenvd(i, j, k1) - fluxse * rsig1[k];
            }
        }
}

/* ---------------------------------------------------------------- sppt.f90 */
bool sppt_on = false;
Spec3 sppt_eta;
static Spec3 sppt_spec;
static FA<double, mx, nx, kx> sppt_sigma;
static bool sppt_first = true;
static const double sppt_mu[kx + 1] = {0, 1, 1, 1, 1, 1, 1, 1, 1};
void sppt_reset() { sppt_first = true; }

/* sppt.f90:45-99.  The reference draws eta from random_number seeded by system_clock
 * (:119-132, not reproducible); the oracle takes the already clipped Gaussian noise
 * eta(m,n,k) from `sppt_eta` so that the CUDA path can be checked on identical noise. */
void gen_sppt(Grid3& sppt_grid) {
    const double time_decorr = 6.0, len_decorr = 500000.0, stddev = F(0.33);
    const double phi_ar = exp(-(24 / (double)nsteps) / time_decorr);   /* :32 24/real(nsteps,p) */
    if (sppt_first) {
        double f0 = 0.0;
        for (int n = 1; n <= trunc_; n++) {
            double r = len_decorr / rearth;
            f0 = f0 + (2 * n + 1) * exp(-0.5 * (r * r) * n * (n + 1));
        }
        f0 = sqrt(((stddev * stddev) * (1 - phi_ar * phi_ar)) / (2 * f0));
        for (int k = 1; k <= kx; k++)
            for (int n = 1; n <= nx; n++)
                for (int m = 1; m <= mx; m++) sppt_sigma(m, n, k) = f0 * exp(-0.25 * (len_decorr * len_decorr) * el2(m, n));
        double c = pow(1 - phi_ar * phi_ar, -0.5);
        for (size_t q = 0; q < sppt_spec.size(); q++) sppt_spec.d[q] = c * sppt_sigma.d[q] * sppt_eta.d[q];
        sppt_first = false;
    } else {
        for (size_t q = 0; q < sppt_spec.size(); q++) sppt_spec.d[q] = phi_ar * sppt_spec.d[q] + sppt_sigma.d[q] * sppt_eta.d[q];
    }
    for (int k = 1; k <= kx; k++) spec_to_grid(sppt_spec.p(1, 1, k), 1, sppt_grid.p(1, 1, k));
    for (size_t q = 0; q < sppt_grid.size(); q++) sppt_grid.d[q] = std::min(1.0, fabs(sppt_grid.d[q])) * copysign(1.0, sppt_grid.d[q]);
}

/* ---------------------------------------------------------------- physics.f90:43-223 */
/* diagnostics kept for the tests (module-level copies of physics' locals) */
FA<int, ix, il> dbg_iptop, dbg_icnv;
FA<int, ix, il, 2> dbg_icltop;
Grid3 dbg_tt_rsw;

void get_physical_tendencies(const cplx* vor_, const cplx* divv, const cplx* tt, const cplx* q, const cplx* phi_, const cplx* psl,
                             Grid3& utend, Grid3& vtend, Grid3& ttend, Grid3& qtend) {
    static Spec2 ucos, vcos;
    static Grid2 pslg, rps, gse, psg, ts, tskin, u0, v0, t0, cloudc, clstr, cltop, prtop;
    static Grid3 ug, vg, tg, qg, phig, utend_dyn, vtend_dyn, ttend_dyn, qtend_dyn, se, rh, qsat;
    static Grid3 tt_cnv, qt_cnv, tt_lsc, qt_lsc, tt_rlw, ut_pbl, vt_pbl, tt_pbl, qt_pbl;
    /* physics.f90:79 tt_rsw is a non-SAVE local that persists between calls in practice (static storage) */
    static Grid3 tt_rsw;
    static FA<int, ix, il> iptop, icnv;
    static FA<int, ix, il, 2> icltop;
    static Grid3 sppt_pattern;
    const size_t sl = (size_t)mx * nx;

    utend_dyn = utend; vtend_dyn = vtend; ttend_dyn = ttend; qtend_dyn = qtend;
    /* :95-104 */
    for (int k = 1; k <= kx; k++) {
        uvspec(vor_ + sl * (k - 1), divv + sl * (k - 1), ucos.p(), vcos.p());
        spec_to_grid(ucos.p(), 2, ug.p(1, 1, k));
        spec_to_grid(vcos.p(), 2, vg.p(1, 1, k));
        spec_to_grid(tt + sl * (k - 1), 1, tg.p(1, 1, k));
        spec_to_grid(q + sl * (k - 1), 1, qg.p(1, 1, k));
        spec_to_grid(phi_ + sl * (k - 1), 1, phig.p(1, 1, k));
    }
    spec_to_grid(psl, 1, pslg.p());
    /* :110-118 */
    GLOOP { psg(i, j) = exp(pslg(i, j)); rps(i, j) = 1.0 / psg(i, j); }
    for (size_t qq = 0; qq < qg.size(); qq++) {
        qg.d[qq] = std::max(qg.d[qq], 0.0);
        se.d[qq] = cp * tg.d[qq] + phig.d[qq];
    }
    for (int k = 1; k <= kx; k++) spec_hum_to_rel_hum(tg.p(1, 1, k), psg.p(), geo.fsg[k], qg.p(1, 1, k), rh.p(1, 1, k), qsat.p(1, 1, k));
    /* :125-138 */
    get_convection_tendencies(psg, se, qg, qsat, iptop, cbmf, precnv, tt_cnv, qt_cnv);
    for (int k = 2; k <= kx; k++)
        GLOOP {
            tt_cnv(i, j, k) = tt_cnv(i, j, k) * rps(i, j) * grdscp[k];
            qt_cnv(i, j, k) = qt_cnv(i, j, k) * rps(i, j) * grdsig[k];
        }
    GLOOP icnv(i, j) = kx - iptop(i, j);
    get_large_scale_condensation_tendencies(psg, qg, qsat, iptop, precls, tt_lsc, qt_lsc);
    for (size_t qq = 0; qq < ttend.size(); qq++) {
        ttend.d[qq] = ttend.d[qq] + tt_cnv.d[qq] + tt_lsc.d[qq];
        qtend.d[qq] = qtend.d[qq] + qt_cnv.d[qq] + qt_lsc.d[qq];
    }
    /* :146-163 */
    if (compute_shortwave) {
        GLOOP gse(i, j) = (se(i, j, kx - 1) - se(i, j, kx)) / (phig(i, j, kx - 1) - phig(i, j, kx));
        clouds(qg, rh, precnv, precls, iptop, gse, fmask_l, icltop, cloudc, clstr);
        IJLOOP {
            cltop(i, j) = sigh[icltop(i, j, 1) - 1] * psg(i, j);
            prtop(i, j) = (double)(float)iptop(i, j);
        }
        get_shortwave_rad_fluxes(psg, qg, icltop, cloudc, clstr, ssrd, ssr, tsr, tt_rsw);
        for (int k = 1; k <= kx; k++)
            GLOOP tt_rsw(i, j, k) = tt_rsw(i, j, k) * rps(i, j) * grdscp[k];
    }
    /* :166-186 */
    get_downward_longwave_rad_fluxes(tg, slrd, tt_rlw);
    get_surface_fluxes(psg, ug, vg, tg, qg, rh, phig, phis0, fmask_l, sst_am, ssrd, slrd, ustr, vstr, shf, evap, slru, hfluxn,
                       ts, tskin, u0, v0, t0, true);
    get_upward_longwave_rad_fluxes(tg, ts, slrd, slru.p(1, 1, 3), slr, olr, tt_rlw);
    for (int k = 1; k <= kx; k++)
        GLOOP tt_rlw(i, j, k) = tt_rlw(i, j, k) * rps(i, j) * grdscp[k];
    for (size_t qq = 0; qq < ttend.size(); qq++) ttend.d[qq] = ttend.d[qq] + tt_rsw.d[qq] + tt_rlw.d[qq];
    /* :193-205 */
    get_vertical_diffusion_tend(se, rh, qg, qsat, phig, icnv, ut_pbl, vt_pbl, tt_pbl, qt_pbl);
    GLOOP {
        ut_pbl(i, j, kx) = ut_pbl(i, j, kx) + ustr(i, j, 3) * rps(i, j) * grdsig[kx];
        vt_pbl(i, j, kx) = vt_pbl(i, j, kx) + vstr(i, j, 3) * rps(i, j) * grdsig[kx];
        tt_pbl(i, j, kx) = tt_pbl(i, j, kx) + shf(i, j, 3) * rps(i, j) * grdscp[kx];
        qt_pbl(i, j, kx) = qt_pbl(i, j, kx) + evap(i, j, 3) * rps(i, j) * grdsig[kx];
    }
    for (size_t qq = 0; qq < ttend.size(); qq++) {
        utend.d[qq] = utend.d[qq] + ut_pbl.d[qq];
        vtend.d[qq] = vtend.d[qq] + vt_pbl.d[qq];
        ttend.d[qq] = ttend.d[qq] + tt_pbl.d[qq];
        qtend.d[qq] = qtend.d[qq] + qt_pbl.d[qq];
    }
    /* :208-222 */
    if (sppt_on) {
        gen_sppt(sppt_pattern);
        for (int k = 1; k <= kx; k++)
            GLOOP {
                double f = (1 + sppt_pattern(i, j, k) * sppt_mu[k]);
                utend(i, j, k) = f * (utend(i, j, k) - utend_dyn(i, j, k)) + utend_dyn(i, j, k);
                vtend(i, j, k) = f * (vtend(i, j, k) - vtend_dyn(i, j, k)) + vtend_dyn(i, j, k);
                ttend(i, j, k) = f * (ttend(i, j, k) - ttend_dyn(i, j, k)) + ttend_dyn(i, j, k);
                qtend(i, j, k) = f * (qtend(i, j, k) - qtend_dyn(i, j, k)) + qtend_dyn(i, j, k);
            }
    }
    dbg_iptop = iptop; dbg_icnv = icnv; dbg_icltop = icltop; dbg_tt_rsw = tt_rsw;
}

}  // namespace orc
