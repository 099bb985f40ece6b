// Model calendar (date.f90) and the per-step flags of the main loop (speedy.f90:27-54), as
// one routine that runs either on the host (plain launches) or in a 1-thread kernel (CUDA
// graph replay, no host involvement).  Integer work is bit-exact; tmonth/tyear are
// evaluated in real32 exactly as date.f90:99-100,150-151.
#pragma once
#include "model.h"

#ifdef __CUDACC__
#define SPD_HD __host__ __device__
#else
#define SPD_HD
#endif

namespace spd {

SPD_HD inline int cal_days_in_month(int m) {   // date.f90:41 ncal365
    const int n[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    return n[m - 1];
}
SPD_HD inline int cal_days_before(int m) {     // ndaycal(m,2), date.f90:91-94
    int s = 0;
    for (int q = 1; q < m; q++) s += cal_days_in_month(q);
    return s;
}

// date.f90:96-100 / :147-151 (iseasc = 1)
SPD_HD inline void cal_fractions(DevClock& c) {
    c.imont1 = c.month;
    c.tmonth = (double)(((float)c.day - 0.5f) / (float)cal_days_in_month(c.month));
    c.tyear = (double)(((float)(cal_days_before(c.month) + c.day) - 0.5f) / 365.0f);
    c.doy = cal_days_before(c.month) + c.day - 1;
    if (c.doy > 364) c.doy = 364;
}

// flags for the step that is about to run
SPD_HD inline void cal_step_flags(DevClock& c) {
    c.do_forcing = ((c.model_step - 1) % c.nsteps == 0) ? 1 : 0;   // speedy.f90:29
    c.csw = (c.model_step % 3 == 1) ? 1 : 0;                 // speedy.f90:35, nstrad = 3
}

// speedy.f90:44-47: model_step += 1; newdate (date.f90:109-157); then the state the coupler
// call of this step sees (couple_sea_atm: obs_ssta on every step of day 1 of a month)
SPD_HD inline void cal_advance(DevClock& c) {
    c.model_step += 1;
    c.minute += 24 * 60 / c.nsteps;                             // date.f90:113
    if (c.minute >= 60) { c.minute = c.minute % 60; c.hour += 1; }
    if (c.hour >= 24) { c.hour = c.hour % 24; c.day += 1; }
    if (c.year % 4 == 0 && c.month == 2) {
        if (c.day > 29) { c.day = 1; c.month += 1; }
    } else {
        if (c.day > cal_days_in_month(c.month)) { c.day = 1; c.month += 1; }
    }
    if (c.month > 12) { c.month = 1; c.year += 1; }
    cal_fractions(c);
    c.obs_ssta = (c.day == 1) ? 1 : 0;                          // sea_model.f90:273 (day argument > 0 in the main loop)
    c.next_month = (c.start_year - 1979) * 12 + c.month;        // sea_model.f90:377, issty0 = 1979
    if (c.obs_ssta && (c.next_month < 1 || c.next_month > c.nssta)) c.ssta_missing = 1;
    cal_step_flags(c);
}

}  // namespace spd
