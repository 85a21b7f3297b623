// Calendar arithmetic on the device: arrow::compute::FloorTemporal / CeilTemporal (no time zone) as
// DataFrame::downsample calls them (/root/reference/src/dataframe.cpp:1265-1290), for every CalendarUnit the
// reference's getCalendarUnit knows (N U L S T H D W M Q Y), any multiple, week_starts_monday and
// calendar_based_origin — including the reference's "subtract one day" step for W / M / Q / Y.
// The integer model is tests/temporal_model.py (checked there against pyarrow's kernels on 6 000 timestamps per
// option combination); this file is its transcription.  Replaces the per-row host relabel of round 1.
#pragma once
#include "common.cuh"

namespace pa {

__device__ __forceinline__ int64_t t_floor_div(int64_t a, int64_t b) {   // b > 0
  int64_t q = a / b;
  return (a % b < 0) ? q - 1 : q;
}

// days since 1970-01-01 <-> proleptic Gregorian civil date (H. Hinnant's algorithms)
__device__ __forceinline__ int64_t t_days_from_civil(int64_t y, int m, int d) {
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const int64_t yoe = y - era * 400;
  const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}
__device__ __forceinline__ void t_civil_from_days(int64_t z, int64_t* y, int* m, int* d) {
  z += 719468;
  const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
  const int64_t doe = z - era * 146097;
  const int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  const int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
  const int64_t mp = (5 * doy + 2) / 153;
  *d = static_cast<int>(doy - (153 * mp + 2) / 5 + 1);
  *m = static_cast<int>(mp + (mp < 10 ? 3 : -9));
  *y = yoe + era * 400 + (*m <= 2);
}

struct TemporalSpec {
  int64_t multiple;       // >= 1
  int unit;               // 'N' 'U' 'L' 'S' 'T' 'H' 'D' 'W' 'M' 'Q' 'Y'
  int ceil;               // CeilTemporal (ceil_is_strictly_greater = false) instead of FloorTemporal
  int week_starts_monday;
  int calendar_origin;    // RoundTemporalOptions::calendar_based_origin
  int64_t ticks_per_sec;  // resolution of the index: 1, 1e3, 1e6, 1e9
  int64_t post_shift;     // added to every label (the reference subtracts one day for W / M / Q / Y)
};

// ticks of one fixed-width unit (0 = finer than the index resolution: rounding is the identity)
__device__ __forceinline__ int64_t t_unit_ticks(int unit, int64_t tps) {
  switch (unit) {
    case 'N': return tps / 1000000000;
    case 'U': return tps / 1000000;
    case 'L': return tps / 1000;
    case 'S': return tps;
    case 'T': return 60 * tps;
    case 'H': return 3600 * tps;
    case 'D': return 86400 * tps;
    default: return 0;
  }
}

__device__ __forceinline__ int64_t t_floor_multiple_epoch(int64_t d, int64_t mult) {   // arrow: d >= 0 ? d / m * m : (d - m + 1) / m * m
  return d >= 0 ? d / mult * mult : (d - mult + 1) / mult * mult;
}

__device__ int64_t t_floor(int64_t t, const TemporalSpec& s) {
  const int64_t tps = s.ticks_per_sec, DAY = 86400 * tps, mult = s.multiple;
  switch (s.unit) {
    case 'N': case 'U': case 'L': case 'S': case 'T': case 'H': case 'D': {
      const int64_t u = t_unit_ticks(s.unit, tps);
      if (u == 0) return t;
      if (mult == 1) return t_floor_div(t, u) * u;
      if (s.calendar_origin) {
        int64_t origin;
        if (s.unit == 'D') {
          int64_t y; int m, d;
          t_civil_from_days(t_floor_div(t, DAY), &y, &m, &d);
          origin = t_days_from_civil(y, m, 1) * DAY;
        } else {
          const int parent = s.unit == 'N' ? 'U' : s.unit == 'U' ? 'L' : s.unit == 'L' ? 'S' : s.unit == 'S' ? 'T' : s.unit == 'T' ? 'H' : 'D';
          const int64_t p = t_unit_ticks(parent, tps);
          origin = t_floor_div(t, p) * p;
        }
        const int64_t mm = mult * u;
        return (t - origin) / mm * mm + origin;
      }
      return t_floor_multiple_epoch(t_floor_div(t, u), mult) * u;
    }
    case 'W': {
      const int64_t off = (s.week_starts_monday ? 3 : 4) * DAY, W = 7 * DAY;
      const int64_t tt = t + off;
      const int64_t d = t_floor_div(tt, W);
      if (mult == 1) return d * W - off;
      if (s.calendar_origin) {
        // weeks counted from the Monday (Sunday) after the last Thursday (Wednesday) of the previous December;
        // arrow does not take the weekday offset off again on this branch
        int64_t y; int m, dd;
        t_civil_from_days(t_floor_div(tt, DAY), &y, &m, &dd);
        const int target = s.week_starts_monday ? 4 : 3;
        const int64_t dec31 = t_days_from_civil(y - 1, 12, 31);
        const int64_t wd31 = ((dec31 + 4) % 7 + 7) % 7;          // 1970-01-01 was a Thursday
        const int64_t last = dec31 - (((wd31 - target) % 7 + 7) % 7);
        const int64_t start = (last + 4) * DAY, uw = mult * W;
        return (tt - start) / uw * uw + start;
      }
      return t_floor_multiple_epoch(d, mult) * W - off;
    }
    case 'M': case 'Q': {
      const int64_t mul = mult * (s.unit == 'Q' ? 3 : 1);
      int64_t y; int m, d;
      t_civil_from_days(t_floor_div(t, DAY), &y, &m, &d);
      if (mul == 1) return t_days_from_civil(y, m, 1) * DAY;
      if (s.calendar_origin) {
        int m0 = m - 1;
        m0 -= static_cast<int>(m0 % mul);
        return t_days_from_civil(y, m0 + 1, 1) * DAY;
      }
      int64_t tm = y * 12 + m - 1 - 1970 * 12;
      tm = t_floor_multiple_epoch(tm, mul);
      const int64_t yy = 1970 + t_floor_div(tm, 12);
      const int mm = static_cast<int>(tm - t_floor_div(tm, 12) * 12);
      return t_days_from_civil(yy, mm + 1, 1) * DAY;
    }
    case 'Y': {
      int64_t y; int m, d;
      t_civil_from_days(t_floor_div(t, DAY), &y, &m, &d);
      return t_days_from_civil(y / mult * mult, 1, 1) * DAY;
    }
    default: return t;
  }
}

__device__ int64_t t_round(int64_t t, const TemporalSpec& s) {
  const int64_t f = t_floor(t, s);
  if (!s.ceil) return f + s.post_shift;
  const int64_t tps = s.ticks_per_sec, DAY = 86400 * tps;
  int64_t c;
  switch (s.unit) {
    case 'W': c = f < t ? f + s.multiple * 7 * DAY : f; break;
    case 'M': case 'Q': {   // arrow: always the NEXT boundary for month / quarter / year
      int64_t y; int m, d;
      t_civil_from_days(t_floor_div(f, DAY), &y, &m, &d);
      const int64_t tm = y * 12 + (m - 1) + s.multiple * (s.unit == 'Q' ? 3 : 1);
      c = t_days_from_civil(t_floor_div(tm, 12), static_cast<int>(tm - t_floor_div(tm, 12) * 12) + 1, 1) * DAY;
      break;
    }
    case 'Y': {
      int64_t y; int m, d;
      t_civil_from_days(t_floor_div(f, DAY), &y, &m, &d);
      c = t_days_from_civil(y + s.multiple, 1, 1) * DAY;
      break;
    }
    default: {
      const int64_t u = t_unit_ticks(s.unit, tps);
      c = (u != 0 && f < t) ? f + s.multiple * u : f;
    }
  }
  return c + s.post_shift;
}

// labels[i] = round(index[i]); null timestamps keep their validity bit (the label value is 0)
__global__ void __launch_bounds__(256) k_temporal_labels(const int64_t* ts, const uint8_t* valid, int64_t bit_off, int64_t n,
                                                         TemporalSpec s, int64_t* out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) out[i] = (!valid || bit_at(valid, bit_off + i)) ? t_round(ts[i], s) : 0;
}

}  // namespace pa
