/* ORACLE (test infrastructure) — restatement of the code that surrounds the hot path and is
 * needed to reach it with the reference's inputs: boundaries.f90, date.f90,
 * interpolation.f90, land_model.f90, sea_model.f90, coupler.f90, forcing.f90,
 * initialization.f90, the main loop of speedy.f90 and the read/flip logic and output
 * conversions of input_output.f90.  Boundary data come from data/bc_t30.bin
 * (tools/pack_boundary.py: the reference's NetCDF variables, unmodified, in file order). */
#include "oracle.h"
#include <map>
#include <string>

namespace orc {

#define GLOOP for (int j = 1; j <= il; j++) for (int i = 1; i <= ix; i++)
#define F(x) ((double)(x##f))

/* ---------------------------------------------------------------- input_output.f90:23-92 */
static std::map<std::string, std::vector<float>> bc_data;   /* name -> nrec*il*ix floats, file order */
static int bc_load(const char* path) {
    bc_data.clear();
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    char magic[8];
    int hdr[3];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "SPDYBC01", 8) != 0 || fread(hdr, 4, 3, f) != 3) { fclose(f); return -2; }
    if (hdr[0] != ix || hdr[1] != il) { fclose(f); return -3; }
    for (int q = 0; q < hdr[2]; q++) {
        char name[17] = {0};
        int nrec;
        if (fread(name, 1, 16, f) != 16 || fread(&nrec, 4, 1, f) != 1) { fclose(f); return -4; }
        std::vector<float> v((size_t)nrec * il * ix);
        if (fread(v.data(), 4, v.size(), f) != v.size()) { fclose(f); return -5; }
        bc_data[name] = std::move(v);
    }
    fclose(f);
    return 0;
}
/* load_boundary_file_*: field = raw_input(:, il:1:-1 [, month]); where (field <= -999) field = 0 */
static void load_boundary_file(const char* field_name, int month, Grid2& field) {
    auto it = bc_data.find(field_name);
    if (it == bc_data.end()) { fprintf(stderr, "oracle: boundary field %s missing\n", field_name); abort(); }
    const size_t nrec = it->second.size() / ((size_t)ix * il);
    if (month < 1 || (size_t)month > nrec) { fprintf(stderr, "oracle: record %d of %s not in the packed file (%zu records)\n", month, field_name, nrec); abort(); }
    const float* raw = it->second.data() + (size_t)(month - 1) * ix * il;
    GLOOP {
        double v = (double)raw[(i - 1) + (size_t)ix * (il - j)];
        if (v <= -999) v = 0.0;
        field(i, j) = v;
    }
}

/* ---------------------------------------------------------------- boundaries.f90 */
Grid2 fmask, phi0, phis0, alb0;

/* boundaries.f90:47-72 */
static void forchk(const Grid2& fm, int nf, double fmin, double fmax, double fset, double* field /*(ix,il,nf)*/) {
    for (int jf = 1; jf <= nf; jf++) {
        int nfault = 0;
        for (int i = 1; i <= ix; i++)
            for (int j = 1; j <= il; j++) {
                double& v = field[(i - 1) + (size_t)ix * ((j - 1) + (size_t)il * (jf - 1))];
                if (fm(i, j) > 0.0) {
                    if (v < fmin || v > fmax) nfault = nfault + 1;
                } else {
                    v = fset;
                }
            }
        (void)nfault;
    }
}
/* boundaries.f90:75-94 */
static void spectral_truncation(Grid2& fg1, Grid2& fg2) {
    Spec2 fsp;
    grid_to_spec(fg1.p(), fsp.p());
    for (int n = 1; n <= nx; n++)
        for (int m = 1; m <= mx; m++)
            if (m + n - 2 > trunc_) fsp(m, n) = cplx(0.0, 0.0);
    spec_to_grid(fsp.p(), 1, fg2.p());
}
/* boundaries.f90:98-142 */
static void fillsf(Grid2& sf, double fmis) {
    double sf2[ix + 2];
    int j1 = 0, j2 = 0, j3 = 0;
    double fmean = 0.0;
    for (int hemisphere = 1; hemisphere <= 2; hemisphere++) {
        if (hemisphere == 1) { j1 = il / 2; j2 = 1; j3 = -1; }
        else { j1 = j1 + 1; j2 = il; j3 = 1; }
        for (int j = j1; (j3 > 0) ? (j <= j2) : (j >= j2); j += j3) {
            for (int i = 1; i <= ix; i++) sf2[i] = sf(i, j);
            int nmis = 0;
            for (int i = 1; i <= ix; i++)
                if (sf(i, j) < fmis) { nmis = nmis + 1; sf2[i] = 0.0; }
            if (nmis < ix) {
                double s = 0.0;
                for (int i = 1; i <= ix; i++) s += sf2[i];
                fmean = s / (double)(float)(ix - nmis);
            }
            for (int i = 1; i <= ix; i++)
                if (sf(i, j) < fmis) sf2[i] = fmean;
            sf2[0] = sf2[ix];
            sf2[ix + 1] = sf2[1];
            for (int i = 1; i <= ix; i++)
                if (sf(i, j) < fmis) sf(i, j) = 0.5 * (sf2[i - 1] + sf2[i + 1]);
        }
    }
}
/* boundaries.f90:28-43 */
static void initialize_boundaries() {
    load_boundary_file("orog", 1, phi0);
    GLOOP phi0(i, j) = grav * phi0(i, j);
    spectral_truncation(phi0, phis0);
    load_boundary_file("lsm", 1, fmask);
    load_boundary_file("alb", 1, alb0);
}

/* ---------------------------------------------------------------- date.f90 */
DateTime model_datetime, start_datetime, end_datetime;
int imont1, isst0;
double tmonth, tyear;
static int ndaycal[13][3];
static const int ncal = 365;
static const int ncal365[13] = {0, 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};

/* date.f90:96-100,147-151: (integer - 0.5) and the quotient are real32 */
static void set_time_fractions() {
    imont1 = model_datetime.month;
    tmonth = (double)(((float)model_datetime.day - 0.5f) / (float)ndaycal[model_datetime.month][1]);
    tyear = (double)(((float)(ndaycal[model_datetime.month][2] + model_datetime.day) - 0.5f) / (float)ncal);
}
/* date.f90:53-105 */
static void initialize_date(int y, int m, int d, int h, int mi) {
    start_datetime = DateTime{y, m, d, h, mi};
    model_datetime = start_datetime;
    for (int jm = 1; jm <= 12; jm++) ndaycal[jm][1] = ncal365[jm];
    ndaycal[1][2] = 0;
    for (int jm = 2; jm <= 12; jm++) ndaycal[jm][2] = ndaycal[jm - 1][1] + ndaycal[jm - 1][2];
    set_time_fractions();
}
/* date.f90:109-157 */
static void newdate() {
    model_datetime.minute = model_datetime.minute + (int)(24 * 60 / nsteps);
    if (model_datetime.minute >= 60) {
        model_datetime.minute = model_datetime.minute % 60;
        model_datetime.hour = model_datetime.hour + 1;
    }
    if (model_datetime.hour >= 24) {
        model_datetime.hour = model_datetime.hour % 24;
        model_datetime.day = model_datetime.day + 1;
    }
    if (model_datetime.year % 4 == 0 && model_datetime.month == 2) {
        if (model_datetime.day > 29) { model_datetime.day = 1; model_datetime.month = model_datetime.month + 1; }
    } else {
        if (model_datetime.day > ndaycal[model_datetime.month][1]) { model_datetime.day = 1; model_datetime.month = model_datetime.month + 1; }
    }
    if (model_datetime.month > 12) { model_datetime.month = 1; model_datetime.year = model_datetime.year + 1; }
    set_time_fractions();
}

void test_calendar_init(int y, int m, int d, int h, int mi) { initialize_date(y, m, d, h, mi); }
void test_newdate() { newdate(); }

/* ---------------------------------------------------------------- interpolation.f90 */
/* :16-35 for12(ix*il,*) */
static void forint(int imon, const double* for12, double* for1) {
    int imon2;
    double wmon;
    if (tmonth <= 0.5) {
        imon2 = imon - 1;
        if (imon == 1) imon2 = 12;
        wmon = 0.5 - tmonth;
    } else {
        imon2 = imon + 1;
        if (imon == 12) imon2 = 1;
        wmon = tmonth - 0.5;
    }
    const size_t N = (size_t)ix * il;
    const double *a = for12 + N * (imon - 1), *b = for12 + N * (imon2 - 1);
    for (size_t q = 0; q < N; q++) for1[q] = a[q] + wmon * (b[q] - a[q]);
}
/* :38-69 */
static void forin5(int imon, const double* for12, double* for1) {
    int im2 = imon - 2, im1 = imon - 1, ip1 = imon + 1, ip2 = imon + 2;
    if (im2 < 1) im2 = im2 + 12;
    if (im1 < 1) im1 = im1 + 12;
    if (ip1 > 12) ip1 = ip1 - 12;
    if (ip2 > 12) ip2 = ip2 - 12;
    double c0 = (double)(1.0f / 12.0f);
    double t0 = c0 * tmonth;
    double t1 = c0 * (1.0 - tmonth);
    double t2 = 0.25 * tmonth * (1 - tmonth);
    double wm2 = -t1 + t2;
    double wm1 = -c0 + 8 * t1 - 6 * t2;
    double w0 = 7 * c0 + 10 * t2;
    double wp1 = -c0 + 8 * t0 - 6 * t2;
    double wp2 = -t0 + t2;
    const size_t N = (size_t)ix * il;
    const double *fm2 = for12 + N * (im2 - 1), *fm1 = for12 + N * (im1 - 1), *f0 = for12 + N * (imon - 1), *fp1 = for12 + N * (ip1 - 1), *fp2 = for12 + N * (ip2 - 1);
    for (size_t q = 0; q < N; q++) for1[q] = wm2 * fm2[q] + wm1 * fm1[q] + w0 * f0[q] + wp1 * fp1[q] + wp2 * fp2[q];
}

/* ---------------------------------------------------------------- land_model.f90 */
static Grid2 rhcapl, cdland, stlcl_ob, snowdcl_ob, soilwcl_ob, bmask_l;
Grid2 stl_am, snowd_am, soilw_am, stl_lm, fmask_l;
static FA<double, ix, il, 12> stl12, snowd12, soilw12;
static const int land_coupling_flag = 1;
static const double sd2sc = 60.0;

/* land_model.f90:50-181 */
static void land_model_init() {
    static Grid2 dmask, veg_low, veg_high, veg, swl1, swl2, tmp;
    const double swcap = F(0.30), swwil = F(0.17), thrsh = F(0.1);
    fmask_l = fmask;
    GLOOP {
        if (fmask_l(i, j) >= thrsh) {
            bmask_l(i, j) = 1.0;
            if (fmask(i, j) > (1.0 - thrsh)) fmask_l(i, j) = 1.0;
        } else {
            bmask_l(i, j) = 0.0;
            fmask_l(i, j) = 0.0;
        }
    }
    for (int month = 1; month <= 12; month++) {
        load_boundary_file("stl", month, tmp);
        fillsf(tmp, 0.0);
        memcpy(stl12.p(1, 1, month), tmp.p(), sizeof(double) * tmp.size());
    }
    forchk(bmask_l, 12, 0.0, 400.0, 273.0, stl12.p());
    for (int month = 1; month <= 12; month++) {
        load_boundary_file("snowd", month, tmp);
        memcpy(snowd12.p(1, 1, month), tmp.p(), sizeof(double) * tmp.size());
    }
    forchk(bmask_l, 12, 0.0, 20000.0, 0.0, snowd12.p());
    load_boundary_file("vegh", 1, veg_high);
    load_boundary_file("vegl", 1, veg_low);
    GLOOP veg(i, j) = std::max(0.0, veg_high(i, j) + F(0.8) * veg_low(i, j));
    const double sdep1 = 70.0;
    const int idep2 = 3;
    const double sdep2 = idep2 * sdep1;
    (void)sdep2;
    const double swwil2 = idep2 * swwil;
    const double rsw = 1.0 / (swcap + idep2 * (swcap - swwil));
    for (int month = 1; month <= 12; month++) {
        load_boundary_file("swl1", month, swl1);
        load_boundary_file("swl2", month, swl2);
        GLOOP {
            double swroot = idep2 * swl2(i, j);
            soilw12(i, j, month) = std::min(1.0, rsw * (swl1(i, j) + veg(i, j) * std::max(0.0, swroot - swwil2)));
        }
    }
    forchk(bmask_l, 12, 0.0, 10.0, 0.0, soilw12.p());
    const double depth_soil = 1.0, depth_lice = 5.0, tdland = 40.;
    const double flandmin = (double)(1.f / 3.f);
    const double hcapl = depth_soil * F(2.50e+6);
    const double hcapli = depth_lice * F(1.93e+6);
    dmask.fill(1.);
    GLOOP if (fmask_l(i, j) < flandmin) dmask(i, j) = 0;
    GLOOP {
        if (alb0(i, j) < F(0.4)) rhcapl(i, j) = delt / hcapl;
        else rhcapl(i, j) = delt / hcapli;
    }
    GLOOP cdland(i, j) = dmask(i, j) * tdland / (1. + dmask(i, j) * tdland);
}
/* land_model.f90:224-239 */
static void run_land_model() {
    GLOOP {
        double tanom = stl_lm(i, j) - stlcl_ob(i, j);
        tanom = cdland(i, j) * (tanom + rhcapl(i, j) * hfluxn(i, j, 1));
        stl_lm(i, j) = tanom + stlcl_ob(i, j);
    }
}
/* land_model.f90:184-221 */
static void couple_land_atm(int day) {
    forin5(imont1, stl12.p(), stlcl_ob.p());
    forint(imont1, snowd12.p(), snowdcl_ob.p());
    forint(imont1, soilw12.p(), soilwcl_ob.p());
    if (day == 0) {
        stl_lm = stlcl_ob;
        stl_am = stlcl_ob;
    } else {
        if (land_coupling_flag == 1) {
            run_land_model();
            stl_am = stl_lm;
        } else {
            stl_am = stlcl_ob;
        }
    }
    snowd_am = snowdcl_ob;
    soilw_am = soilwcl_ob;
}

/* ---------------------------------------------------------------- sea_model.f90 */
static Grid2 rhcaps, rhcapi, cdsea, cdice, bmask_s, hfseacl, sicecl_ob, ticecl_ob, sstan_ob, sstan_am, wsst_ob;
static double deglat_s[il + 1];
static FA<double, ix, il, 12> sst12, sice12;
static FA<double, ix, il, 3> sstan3;
Grid2 fmask_s, sstcl_ob, sst_am, sice_am, tice_am, ssti_om, sst_om, tice_om, sice_om;
int sea_coupling_flag = 0;
static const int ice_coupling_flag = 1, sst_anomaly_coupling_flag = 1;
static double beta_ = 1.0;

/* sea_model.f90:80-250 */
static void sea_model_init() {
    static Grid2 dmask, tmp;
    double hcaps[il + 1], hcapi[il + 1];
    const double depth_ml = 60., dept0_ml = 40., depth_ice = 2.5, dept0_ice = 1.5, tdsst = 90., tdice = 30.0;
    const double fseamin = (double)(1.f / 3.f), thrsh = F(0.1);
    GLOOP {
        fmask_s(i, j) = 1.0 - fmask(i, j);
        if (fmask_s(i, j) >= thrsh) {
            bmask_s(i, j) = 1.0;
            if (fmask_s(i, j) > (1.0 - thrsh)) fmask_s(i, j) = 1.0;
        } else {
            bmask_s(i, j) = 0.0;
            fmask_s(i, j) = 0.0;
        }
    }
    for (int j = 1; j <= il; j++) deglat_s[j] = geo.radang[j] * 90.0 / (double)asinf(1.0f);   /* :153 */
    for (int month = 1; month <= 12; month++) {
        load_boundary_file("sst", month, tmp);
        fillsf(tmp, 0.0);
        memcpy(sst12.p(1, 1, month), tmp.p(), sizeof(double) * tmp.size());
    }
    forchk(bmask_s, 12, 100.0, 400.0, 273.0, sst12.p());
    for (int month = 1; month <= 12; month++) {
        load_boundary_file("icec", month, tmp);
        GLOOP sice12(i, j, month) = std::max(tmp(i, j), 0.0);
    }
    forchk(bmask_s, 12, 0.0, 1.0, 0.0, sice12.p());
    if (sst_anomaly_coupling_flag > 0) {
        for (int month = 1; month <= 3; month++)
            if ((isst0 <= 1 && month != 2) || isst0 > 1) {
                load_boundary_file("ssta", isst0 - 2 + month, tmp);
                memcpy(sstan3.p(1, 1, month), tmp.p(), sizeof(double) * tmp.size());
            }
        forchk(bmask_s, 3, -50.0, 50.0, 0.0, sstan3.p());
    }
    hfseacl.fill(0.0);
    beta_ = 1.;
    const double crad = (double)(asinf(1.f) / 90.f);   /* :208 all real32 */
    for (int j = 1; j <= il; j++) {
        double coslat = cos(crad * deglat_s[j]);
        hcaps[j] = F(4.18e+6) * (depth_ml + (dept0_ml - depth_ml) * (coslat * coslat * coslat));
        hcapi[j] = F(1.93e+6) * (depth_ice + (dept0_ice - depth_ice) * (coslat * coslat));
    }
    dmask.fill(1.);   /* l_globe */
    for (int j = 2; j <= il - 1; j++)
        for (int i = 1; i <= ix; i++) rhcaps(i, j) = 0.25 * (dmask(i, j - 1) + 2 * dmask(i, j) + dmask(i, j + 1));
    for (int j = 2; j <= il - 1; j++)
        for (int i = 1; i <= ix; i++) dmask(i, j) = rhcaps(i, j);
    GLOOP if (fmask_s(i, j) < fseamin) dmask(i, j) = 0;
    GLOOP { rhcaps(i, j) = delt / hcaps[j]; rhcapi(i, j) = delt / hcapi[j]; }
    GLOOP {
        cdsea(i, j) = dmask(i, j) * tdsst / (1. + dmask(i, j) * tdsst);
        cdice(i, j) = dmask(i, j) * tdice / (1. + dmask(i, j) * tdice);
    }
}
/* sea_model.f90:366-385 */
static void obs_ssta() {
    static Grid2 tmp;
    GLOOP { sstan3(i, j, 1) = sstan3(i, j, 2); sstan3(i, j, 2) = sstan3(i, j, 3); }
    int next_month = (start_datetime.year - issty0) * 12 + model_datetime.month;
    load_boundary_file("ssta", next_month, tmp);
    memcpy(sstan3.p(1, 1, 3), tmp.p(), sizeof(double) * tmp.size());
    forchk(bmask_s, 1, -50.0, 50.0, 0.0, sstan3.p(1, 1, 3));
}
/* sea_model.f90:387-444 */
static void run_sea_model() {
    const double sstfr = (double)(273.2f - 1.8f);   /* :400 real32 subtraction */
    const double sstfr4 = (sstfr * sstfr) * (sstfr * sstfr);
    GLOOP {
        double ti = tice_am(i, j);
        double difice = (albsea - albice) * ssrd(i, j) + emisfc * sbc * (sstfr4 - (ti * ti) * (ti * ti)) + shf(i, j, 2) + evap(i, j, 2) * alhc;
        double hflux_i = hfluxn(i, j, 2) + difice * (1.0 - sice_am(i, j));
        double hflux = hfluxn(i, j, 2) - hfseacl(i, j) - sicecl_ob(i, j) * (hflux_i + beta_ * (sstfr - tice_om(i, j)));
        double tanom = sst_om(i, j) - sstcl_ob(i, j);
        tanom = cdsea(i, j) * (tanom + rhcaps(i, j) * hflux);
        sst_om(i, j) = tanom + sstcl_ob(i, j);
        hflux = hflux_i + beta_ * (sstfr - tice_om(i, j));
        tanom = tice_om(i, j) - ticecl_ob(i, j);
        const double anom0 = 20.;
        double cdis = cdice(i, j) * (anom0 / (anom0 + fabs(tanom)));
        tanom = cdis * (tanom + rhcapi(i, j) * hflux);
        tice_om(i, j) = tanom + ticecl_ob(i, j);
        sice_om(i, j) = sicecl_ob(i, j);
    }
}
/* sea_model.f90:253-363 */
static void couple_sea_atm(int day) {
    forin5(imont1, sst12.p(), sstcl_ob.p());
    forint(imont1, sice12.p(), sicecl_ob.p());
    if (sst_anomaly_coupling_flag > 0) {
        if (model_datetime.day == 1 && day > 0) obs_ssta();
        forint(2, sstan3.p(), sstan_ob.p());
    }
    const double sstfr = (double)(273.2f - 1.8f);   /* :285 */
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            if (sstcl_ob(i, j) > sstfr) {
                sicecl_ob(i, j) = std::min(0.5, sicecl_ob(i, j));
                ticecl_ob(i, j) = sstfr;
                if (sicecl_ob(i, j) > 0.0) sstcl_ob(i, j) = sstfr + (sstcl_ob(i, j) - sstfr) / (1.0 - sicecl_ob(i, j));
            } else {
                sicecl_ob(i, j) = std::max(0.5, sicecl_ob(i, j));
                ticecl_ob(i, j) = sstfr + (sstcl_ob(i, j) - sstfr) / sicecl_ob(i, j);
                sstcl_ob(i, j) = sstfr;
            }
        }
    if (day == 0) {
        sst_om = sstcl_ob;
        tice_om = ticecl_ob;
        sice_om = sicecl_ob;
        if (sea_coupling_flag <= 0) sst_om.fill(0.0);
        wsst_ob.fill(0.);
    } else {
        if (sea_coupling_flag > 0 || ice_coupling_flag > 0) run_sea_model();
    }
    sstan_am.fill(0.0);
    if (sea_coupling_flag <= 1) {
        if (sst_anomaly_coupling_flag > 0) sstan_am = sstan_ob;
        GLOOP sst_am(i, j) = sstcl_ob(i, j) + sstan_am(i, j);
    }
    if (ice_coupling_flag > 0) {
        sice_am = sice_om;
        tice_am = tice_om;
    } else {
        sice_am = sicecl_ob;
        tice_am = ticecl_ob;
    }
    GLOOP {
        sst_am(i, j) = sst_am(i, j) + sice_am(i, j) * (tice_am(i, j) - sst_am(i, j));
        ssti_om(i, j) = sst_om(i, j) + sice_am(i, j) * (tice_am(i, j) - sst_om(i, j));
    }
}

/* ---------------------------------------------------------------- coupler.f90 */
static void initialize_coupler() {
    land_model_init();
    couple_land_atm(0);
    sea_model_init();
    couple_sea_atm(0);
}
static void couple_sea_land(int day) {
    couple_land_atm(day);
    couple_sea_atm(day);
}

/* ---------------------------------------------------------------- forcing.f90:15-116 */
static void set_forcing(int imode) {
    static Grid2 corh, tsfc, tref_, psfc, qsfc, qref, ones;
    double gamlat[il + 1];
    if (imode == 0) {
        radset();
        set_orog_land_sfc_drag(phis0);
        ablco2_ref = ablco2;
    }
    get_zonal_average_fields(tyear);
    for (int i = 1; i <= ix; i++)
        for (int j = 1; j <= il; j++) {
            snowc(i, j) = std::min(1.0, snowd_am(i, j) / sd2sc);
            alb_l(i, j) = alb0(i, j) + snowc(i, j) * (albsn - alb0(i, j));
            alb_s(i, j) = albsea + sice_am(i, j) * (albice - albsea);
            albsfc(i, j) = alb_s(i, j) + fmask_l(i, j) * (alb_l(i, j) - alb_s(i, j));
        }
    /* setgam :104-116 */
    gamlat[1] = gamma_ / (1000. * grav);
    for (int j = 2; j <= il; j++) gamlat[j] = gamlat[1];
    GLOOP corh(i, j) = gamlat[j] * phis0(i, j);
    grid_to_spec(corh.p(), tcorh.p());
    for (int j = 1; j <= il; j++) {
        double pexp = 1. / (rgas * gamlat[j]);
        for (int i = 1; i <= ix; i++) {
            tsfc(i, j) = fmask_l(i, j) * stl_am(i, j) + fmask_s(i, j) * sst_am(i, j);
            tref_(i, j) = tsfc(i, j) + corh(i, j);
            psfc(i, j) = pow(tsfc(i, j) / tref_(i, j), pexp);
        }
    }
    GLOOP ones(i, j) = psfc(i, j) / psfc(i, j);
    get_qsat(tref_.p(), ones.p(), -1.0, qref.p());
    get_qsat(tsfc.p(), psfc.p(), 1.0, qsfc.p());
    GLOOP corh(i, j) = refrh1 * (qref(i, j) - qsfc(i, j));
    grid_to_spec(corh.p(), qcorh.p());
}

/* ---------------------------------------------------------------- initialization.f90:12-82 + speedy.f90 */
int model_step = 1;

int model_initialize(const char* bc_file, int y, int m, int d, int h, int mi) {
    int rc = bc_load(bc_file);
    if (rc) return rc;
    initialize_date(y, m, d, h, mi);
    isst0 = (start_datetime.year - issty0) * 12 + start_datetime.month;
    initialize_geometry();
    initialize_spectral();
    initialize_geopotential();
    initialize_horizontal_diffusion();
    initialize_physics();
    initialize_boundaries();
    /* module state that the reference initialises statically */
    compute_shortwave = true;
    ablco2 = 6.0;
    sppt_reset();
    initialize_prognostics();
    initialize_coupler();
    set_forcing(0);
    first_step();
    model_step = 1;
    return 0;
}

/* speedy.f90:27-54 loop body, repeated */
int model_run_steps(int nrun) {
    double diag[kx * 3];
    for (int s = 0; s < nrun; s++) {
        if ((model_step - 1) % nsteps == 0) set_forcing(1);
        compute_shortwave = (model_step % nstrad) == 1;
        step(2, 2, 2 * delt);
        if (check_diagnostics(vor.p(1, 1, 1, 2), div_.p(1, 1, 1, 2), t.p(1, 1, 1, 2), model_step, diag, false)) return 1;
        model_step = model_step + 1;
        newdate();
        couple_sea_land(1 + model_step / nsteps);
    }
    return 0;
}

}  // namespace orc
