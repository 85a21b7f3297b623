// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement of PandasArrow's group-by / resample hot path, expressed as the same
// sequence of Apache Arrow C++ calls the reference makes.  The reference itself cannot be
// compiled in this image (its headers need Boost.date_time, oneTBB, range-v3, spdlog,
// tabulate, hosseinmoein/DataFrame and a private cmake include), but all of
// the arithmetic on this path lives in Arrow C++, and pyarrow 24.0.0 ships the headers and
// libarrow/libarrow_compute.  Pin: Arrow 24.0.0 (the reference's CMake does not pin one).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this library.  The product (pandasarrow_b200/csrc) never links or calls it.
//
// Parity pinning: tests/test_oracle_golden.py checks this file against every known-answer
// vector the reference's own tests hold for the path.
//
// Reference call sites restated (paths relative to /root/reference/src):
//   makeGroups            dataframe.cpp:1571-1600  -> Groups::build()
//   processIndex/Each     dataframe.cpp:1539-1569  -> Groups::gather_column()
//   GROUPBY_AGG           pd_core_macros.h:80-147  -> agg_boxed()   (sum/min/max/product)
//   GROUPBY_NUMERIC_AGG   pd_core_macros.h:5-78    -> agg_numeric() (mean/count; validity dropped)
//   GroupBy::first/last   dataframe.cpp:1698-1806  -> agg_position()
//   GroupBy::min_max      dataframe.cpp:1602-1696  -> agg_min_max()
//   HashScalar            ndframe.h:20-37          -> ScalarKeyHash / ScalarKeyEq
//   makeGroupInfo & co.   resample.cpp:11-295, resample.h:9-122 -> resample_labels()
//   DataFrame::downsample dataframe.cpp:1265-1290  -> downsample_labels()
//   NDFrame aggregations  ndframe.cpp:26-55,119-241 -> orc_scalar_agg()
// TBB parallel_for over groups is replaced by OpenMP (TBB is not installed).

#include <arrow/api.h>
#include <arrow/c/bridge.h>
#include <arrow/compute/api.h>
#include <arrow/compute/row/grouper.h>

#include <chrono>
#include <cstring>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace ac = arrow::compute;
using arrow::Array;
using arrow::ArrayVector;
using arrow::Datum;
using arrow::Result;
using arrow::Scalar;
using arrow::Status;

namespace {

thread_local std::string g_err;

// group_by.h:219 — `defaultOpt` is an empty shared_ptr<FunctionOptions>; `.get()` is a typed null.
const ac::FunctionOptions* const kDefaultOptions = nullptr;

int fail(const Status& st) {
  g_err = st.ToString();
  return 1;
}
int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

double now_ms() {
  using clk = std::chrono::steady_clock;
  return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}

void ensure_compute_initialized() {
  static const bool once = [] {
    auto st = ac::Initialize();
    (void)st;
    return true;
  }();
  (void)once;
}

// ndframe.h:20-37 — the map key is a boxed scalar; hash = Scalar::hash(), equality casts the
// right-hand side to the left-hand type before Equals().
struct ScalarKeyHash {
  size_t operator()(const std::shared_ptr<Scalar>& s) const { return s->hash(); }
};
struct ScalarKeyEq {
  bool operator()(const std::shared_ptr<Scalar>& a, const std::shared_ptr<Scalar>& b) const {
    auto cast = b->CastTo(a->type);
    return cast.ok() && a->Equals(**cast);
  }
};
using GroupSlices =
    std::unordered_map<std::shared_ptr<Scalar>, ArrayVector, ScalarKeyHash, ScalarKeyEq>;
using IndexSlices = std::unordered_map<std::shared_ptr<Scalar>, std::shared_ptr<Array>,
                                       ScalarKeyHash, ScalarKeyEq>;

struct Timing {
  double consume_ms = 0, groupings_ms = 0, gather_ms = 0, total_ms = 0;
};

struct Groups {
  std::shared_ptr<arrow::RecordBatch> frame;   // df.m_array
  std::shared_ptr<Array> index;                // df.m_index (may be null -> 0..N-1)
  std::vector<std::string> key_names;
  std::unique_ptr<ac::Grouper> grouper;
  std::shared_ptr<arrow::UInt32Array> row_ids;   // Consume() output
  std::shared_ptr<arrow::ListArray> groupings;   // MakeGroupings() output
  ArrayVector unique_keys;                       // one array per key column
  // Reference-faithful materialisation: every column gathered into group order, sliced per
  // group, stored under the boxed key scalar (single-key mode only).
  GroupSlices slices;
  IndexSlices index_slices;
  // Gathered list arrays per column (index by schema position); filled lazily when
  // materialize == 0 so that large parity cases do not pay for unused columns.
  std::vector<std::shared_ptr<arrow::ListArray>> gathered;
  bool materialized = false;
  Timing timing;

  int64_t num_groups() const { return grouper ? grouper->num_groups() : 0; }

  Status build(bool materialize) {
    const double t0 = now_ms();
    std::vector<Datum> key_data;
    for (const auto& name : key_names) {
      if (name == "__resampler_idx__") {
        if (!index) return Status::Invalid("frame has no index");
        key_data.emplace_back(index);
      } else {
        auto col = frame->GetColumnByName(name);
        if (!col) return Status::KeyError("no such column: ", name);
        key_data.emplace_back(col);
      }
    }
    ARROW_ASSIGN_OR_RAISE(auto key_batch, ac::ExecBatch::Make(key_data));
    ARROW_ASSIGN_OR_RAISE(grouper, ac::Grouper::Make(key_batch.GetTypes()));
    const double t1 = now_ms();
    ARROW_ASSIGN_OR_RAISE(Datum ids, grouper->Consume(ac::ExecSpan(key_batch)));
    row_ids = ids.array_as<arrow::UInt32Array>();
    const double t2 = now_ms();
    ARROW_ASSIGN_OR_RAISE(groupings,
                          ac::Grouper::MakeGroupings(*row_ids, grouper->num_groups()));
    ARROW_ASSIGN_OR_RAISE(auto uniques, grouper->GetUniques());
    for (auto& v : uniques.values) unique_keys.push_back(v.make_array());
    const double t3 = now_ms();
    gathered.assign(frame->num_columns(), nullptr);
    if (materialize) {
      if (unique_keys.size() != 1)
        return Status::Invalid("materialize=1 follows the reference API: exactly one key");
      if (index) {
        ARROW_ASSIGN_OR_RAISE(auto g, ac::Grouper::ApplyGroupings(*groupings, *index));
        for (int64_t i = 0; i < num_groups(); ++i) {
          ARROW_ASSIGN_OR_RAISE(auto key, unique_keys[0]->GetScalar(i));
          index_slices[key] = g->value_slice(i);
        }
      }
      for (int c = 0; c < frame->num_columns(); ++c) {
        ARROW_ASSIGN_OR_RAISE(gathered[c],
                              ac::Grouper::ApplyGroupings(*groupings, *frame->column(c)));
        for (int64_t i = 0; i < num_groups(); ++i) {
          ARROW_ASSIGN_OR_RAISE(auto key, unique_keys[0]->GetScalar(i));
          slices[key].emplace_back(gathered[c]->value_slice(i));
        }
      }
      materialized = true;
    }
    const double t4 = now_ms();
    timing.consume_ms = t2 - t1;
    timing.groupings_ms = t3 - t2;
    timing.gather_ms = t4 - t3;
    timing.total_ms = t4 - t0;
    return Status::OK();
  }

  Result<int> column_index(const std::string& name) const {
    int idx = frame->schema()->GetFieldIndex(name);
    if (idx < 0) return Status::KeyError("no such column: ", name);
    return idx;
  }

  // The j-th group's slice of column c, through the same lookup the reference performs
  // (boxed key -> map -> vector[c]) when materialised, else straight from the list array.
  Result<std::shared_ptr<Array>> group_slice(int c, int64_t j) {
    if (materialized) {
      ARROW_ASSIGN_OR_RAISE(auto key, unique_keys[0]->GetScalar(j));
      return slices.at(key)[c];
    }
    return gathered[c]->value_slice(j);
  }

  Status ensure_gathered(int c) {
    if (!gathered[c]) {
      ARROW_ASSIGN_OR_RAISE(gathered[c],
                            ac::Grouper::ApplyGroupings(*groupings, *frame->column(c)));
    }
    return Status::OK();
  }
};

Result<std::shared_ptr<Array>> scalars_to_array(const arrow::ScalarVector& v,
                                                const std::shared_ptr<arrow::DataType>& ty) {
  // group_by.h:191-217 builds from v.back()->type; for G == 0 the reference dereferences a
  // null builder — the oracle returns an empty array of the expected type instead.
  std::unique_ptr<arrow::ArrayBuilder> builder;
  ARROW_RETURN_NOT_OK(arrow::MakeBuilder(arrow::default_memory_pool(),
                                         v.empty() ? ty : v.back()->type, &builder));
  if (!v.empty()) ARROW_RETURN_NOT_OK(builder->AppendScalars(v));
  return builder->Finish();
}

// pd_core_macros.h:114-147 — one CallFunction(name, {slice}) per group with default options.
Result<std::shared_ptr<Array>> agg_boxed(Groups& g, const std::string& func,
                                         const std::string& column, int nthreads) {
  ARROW_ASSIGN_OR_RAISE(int c, g.column_index(column));
  ARROW_RETURN_NOT_OK(g.ensure_gathered(c));
  const int64_t G = g.num_groups();
  arrow::ScalarVector out(G);
  Status first_error;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t j = 0; j < G; ++j) {
    auto slice = g.group_slice(c, j);
    if (!slice.ok()) {
#pragma omp critical
      first_error = slice.status();
      continue;
    }
    auto d = ac::CallFunction(func, {*slice}, kDefaultOptions);
    if (!d.ok()) {
#pragma omp critical
      first_error = d.status();
      continue;
    }
    out[j] = d->scalar();
  }
  ARROW_RETURN_NOT_OK(first_error);
  ARROW_ASSIGN_OR_RAISE(
      auto probe, ac::CallFunction(func, {g.frame->column(c)->Slice(0, 0)}, kDefaultOptions));
  return scalars_to_array(out, probe.scalar()->type);
}

// pd_core_macros.h:47-78 — same loop, but the typed scalar's `.value` member is copied into a
// std::vector<T>, so an all-null group's null result turns into whatever `.value` holds
// (validity is dropped).  `valid_out`, when non-null, additionally records the scalar's
// validity so that tests can tell the two apart.
template <typename CType, typename ScalarType, typename BuilderType>
Result<std::shared_ptr<Array>> agg_numeric_typed(Groups& g, const std::string& func, int c,
                                                 int nthreads,
                                                 std::shared_ptr<Array>* valid_out) {
  const int64_t G = g.num_groups();
  std::vector<CType> out(G);
  std::vector<uint8_t> valid(G, 1);
  Status first_error;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t j = 0; j < G; ++j) {
    auto slice = g.group_slice(c, j);
    if (!slice.ok()) {
#pragma omp critical
      first_error = slice.status();
      continue;
    }
    auto d = ac::CallFunction(func, {*slice}, kDefaultOptions);
    if (!d.ok()) {
#pragma omp critical
      first_error = d.status();
      continue;
    }
    const auto& s = d->template scalar_as<ScalarType>();
    out[j] = s.value;
    valid[j] = s.is_valid;
  }
  ARROW_RETURN_NOT_OK(first_error);
  BuilderType b;
  ARROW_RETURN_NOT_OK(b.AppendValues(out));
  if (valid_out) {
    arrow::BooleanBuilder vb;
    ARROW_RETURN_NOT_OK(vb.AppendValues(valid));
    ARROW_ASSIGN_OR_RAISE(*valid_out, vb.Finish());
  }
  return b.Finish();
}

Result<std::shared_ptr<Array>> agg_numeric(Groups& g, const std::string& func,
                                           const std::string& column, int nthreads,
                                           std::shared_ptr<Array>* valid_out) {
  ARROW_ASSIGN_OR_RAISE(int c, g.column_index(column));
  ARROW_RETURN_NOT_OK(g.ensure_gathered(c));
  if (func == "mean" || func == "variance" || func == "stddev")  // dataframe.cpp:1512,1516,1520
    return agg_numeric_typed<double, arrow::DoubleScalar, arrow::DoubleBuilder>(g, func, c,
                                                                                nthreads, valid_out);
  if (func == "count" || func == "count_distinct")  // dataframe.cpp:1526,1528
    return agg_numeric_typed<int64_t, arrow::Int64Scalar, arrow::Int64Builder>(g, func, c,
                                                                               nthreads, valid_out);
  if (func == "all" || func == "any")  // dataframe.cpp:1522,1524 (std::vector<bool> there; bytes here, same values)
    return agg_numeric_typed<uint8_t, arrow::BooleanScalar, arrow::BooleanBuilder>(g, func, c, nthreads, valid_out);
  return Status::NotImplemented("numeric aggregate ", func);
}

// dataframe.cpp:1730-1749 / 1784-1806 — positional first/last (nulls are not skipped).
Result<std::shared_ptr<Array>> agg_position(Groups& g, bool last, const std::string& column) {
  ARROW_ASSIGN_OR_RAISE(int c, g.column_index(column));
  ARROW_RETURN_NOT_OK(g.ensure_gathered(c));
  const int64_t G = g.num_groups();
  arrow::ScalarVector out(G);
  for (int64_t j = 0; j < G; ++j) {
    ARROW_ASSIGN_OR_RAISE(auto slice, g.group_slice(c, j));
    ARROW_ASSIGN_OR_RAISE(out[j], slice->GetScalar(last ? slice->length() - 1 : 0));
  }
  return scalars_to_array(out, g.frame->column(c)->type());
}

// dataframe.cpp:1655-1696 — arrow::compute::MinMax per group; struct {min,max} split in two.
Status agg_min_max(Groups& g, const std::string& column, int nthreads,
                   std::shared_ptr<Array>* mn, std::shared_ptr<Array>* mx) {
  ARROW_ASSIGN_OR_RAISE(int c, g.column_index(column));
  ARROW_RETURN_NOT_OK(g.ensure_gathered(c));
  const int64_t G = g.num_groups();
  arrow::ScalarVector lo(G), hi(G);
  Status first_error;
#pragma omp parallel for schedule(static) num_threads(nthreads)
  for (int64_t j = 0; j < G; ++j) {
    auto slice = g.group_slice(c, j);
    if (!slice.ok()) {
#pragma omp critical
      first_error = slice.status();
      continue;
    }
    auto d = ac::MinMax(*slice);
    if (!d.ok()) {
#pragma omp critical
      first_error = d.status();
      continue;
    }
    const auto& st = d->scalar_as<arrow::StructScalar>();
    lo[j] = st.value[0];
    hi[j] = st.value[1];
  }
  ARROW_RETURN_NOT_OK(first_error);
  ARROW_ASSIGN_OR_RAISE(*mn, scalars_to_array(lo, g.frame->column(c)->type()));
  ARROW_ASSIGN_OR_RAISE(*mx, scalars_to_array(hi, g.frame->column(c)->type()));
  return Status::OK();
}

// ---------------------------------------------------------------------------------------
// Resample (resample.cpp / resample.h).  Boost ptime/time_duration are replaced by int64
// nanoseconds since the Unix epoch; fixed-width rules (time_duration) first, the DateOffset branch
//  further down on std::chrono's civil calendar.
// ---------------------------------------------------------------------------------------
constexpr int64_t kNsPerDay = 86400LL * 1000000000LL;

int64_t floor_to_day(int64_t ns) {
  int64_t d = ns / kNsPerDay;
  if (ns % kNsPerDay < 0) --d;
  return d * kNsPerDay;
}

enum Origin { kEpoch = 0, kStart = 1, kStartDay = 2, kEnd = 3, kEndDay = 4, kCustom = 5 };

// resample.cpp:85-178 (tz handling, :126-136, is not restated: tz must be empty).
void adjust_dates_anchored(int64_t start, int64_t end, int64_t freq, bool closed_right,
                           int origin, int64_t origin_custom, int64_t offset, int64_t* first,
                           int64_t* last) {
  int64_t f = start, l = end;
  int64_t origin_ns = 0;
  switch (origin) {
    case kStartDay: origin_ns = floor_to_day(f); break;
    case kStart: origin_ns = f; break;
    case kEnd: origin_ns = l; break;
    case kEndDay: origin_ns = floor_to_day(l); break;  // ptime(last.date()) — midnight of last day
    case kCustom: origin_ns = origin_custom; break;
    default: origin_ns = 0;
  }
  origin_ns += offset;
  const int64_t foffset = (f - origin_ns) % freq;  // C++ truncating %, as total_nanoseconds() % ...
  const int64_t loffset = (l - origin_ns) % freq;
  if (closed_right) {
    if (foffset > 0) f -= foffset; else f -= freq;
    if (loffset > 0) l += (freq - loffset);
  } else {
    if (foffset > 0) f -= foffset;
    if (loffset > 0) l += (freq - loffset); else l += freq;
  }
  *first = f;
  *last = l;
}

// resample.cpp:11-83 restricted to null-free input (the reference's null branch re-reads
// null_count() after DropNull, which is always 0, so the shift at :72-80 never fires).
Status generate_bins(const int64_t* values, int64_t n, const std::vector<int64_t>& edges,
                     bool right_closed, std::vector<int64_t>* bins) {
  const int64_t nb = static_cast<int64_t>(edges.size());
  if (n <= 0 || nb <= 0) return Status::Invalid("Invalid length for values or for binner");
  if (values[0] < edges[0]) return Status::Invalid("Values falls before first bin");
  if (values[n - 1] > edges[nb - 1]) return Status::Invalid("Values falls after last bin");
  bins->clear();
  bins->reserve(nb);
  int64_t j = 0;
  for (int64_t i = 0; i + 1 < nb; ++i) {
    const int64_t r = edges[i + 1];
    if (right_closed) while (j < n && values[j] <= r) ++j;
    else              while (j < n && values[j] < r) ++j;
    bins->push_back(j);
  }
  return Status::OK();
}

// makeGroupInfo (resample.cpp:202-295) for a time_duration rule + GroupInfo::downsample
// (resample.h:19-43): returns the per-row label array that becomes the new index.
Result<std::shared_ptr<Array>> resample_labels(const std::shared_ptr<Array>& ax, int64_t freq_ns,
                                               bool closed_right, bool label_right, int origin,
                                               int64_t origin_custom, int64_t offset_ns) {
  auto ts = std::dynamic_pointer_cast<arrow::TimestampArray>(ax);
  if (!ts) return Status::Invalid("axis must be a TimestampArray but got array of type ",
                                  ax->type()->ToString());
  if (ts->null_count() > 0) return Status::NotImplemented("oracle: null timestamps");
  if (freq_ns <= 0) return Status::Invalid("FREQ must be positive");
  arrow::TimestampBuilder out(ts->type(), arrow::default_memory_pool());
  if (ts->length() == 0) return out.Finish();
  ARROW_ASSIGN_OR_RAISE(Datum mm, ac::MinMax(ts));
  const auto& st = mm.scalar_as<arrow::StructScalar>();
  const int64_t mn = static_cast<const arrow::TimestampScalar&>(*st.value[0]).value;
  const int64_t mx = static_cast<const arrow::TimestampScalar&>(*st.value[1]).value;
  int64_t first, last;
  adjust_dates_anchored(mn, mx, freq_ns, closed_right, origin, origin_custom, offset_ns, &first,
                        &last);
  if (first >= last) return Status::Invalid("start date has to be less than end date");
  std::vector<int64_t> binner;  // core.cpp:308-331 date_range(first, last, freq)
  for (int64_t t = first; t <= last; t += freq_ns) binner.push_back(t);
  std::vector<int64_t> bins;
  ARROW_RETURN_NOT_OK(generate_bins(ts->raw_values(), ts->length(), binner, closed_right, &bins));
  // label selection, resample.cpp:269-292
  size_t label_begin = label_right ? 1 : 0;
  size_t n_labels = binner.size() - label_begin;
  if (bins.size() < n_labels) n_labels = bins.size();
  // resample.h:14-17,102-105
  if (bins.back() < static_cast<int64_t>(n_labels))
    return Status::NotImplemented("upSampling is not implemented.");
  if (bins.size() != n_labels)
    return Status::Invalid("Processing Group Info requires bins.size() == labels->length()");
  std::vector<int64_t> labels(bins.back());
  int64_t prev = 0;
  for (size_t b = 0; b < bins.size(); ++b) {
    for (int64_t i = prev; i < bins[b]; ++i) labels[i] = binner[label_begin + b];
    prev = bins[b];
  }
  ARROW_RETURN_NOT_OK(out.AppendValues(labels));
  return out.Finish();
}

// makeGroupInfo's DateOffset branch (resample.cpp:248-267) + GroupInfo::downsample.  Boost.date_time is replaced by
// std::chrono's civil calendar (sys_days / year_month_day): DateOffset::add (core.cpp:12-60), the day / week / month /
// year iterators of date_range (core.cpp:175-265) and adjustBinEdges (resample.cpp:180-200) are restated on it.
// Offset types are the reference's DateOffset::Type enumerators (core.h:122-134).
enum OffsetType { kDay = 0, kMonthEnd = 1, kQuarterStart = 2, kQuarterEnd = 3, kWeekStart = 4, kWeekEnd = 5,
                  kMonthStart = 6, kYearEnd = 7, kYearStart = 8 };

namespace cal = std::chrono;

// core.cpp:12-60 for the types date_range accepts.  `currentDate += months(k)` only contributes its year / month:
// every one of these branches then rebuilds the date on day 1.
cal::sys_days offset_add(cal::sys_days d, int type, int k) {
  const cal::year_month_day ymd{d};
  switch (type) {
    case kMonthStart: {
      const cal::year_month ym = cal::year_month{ymd.year(), ymd.month()} + cal::months{k};
      return cal::sys_days{ym / cal::day{1}};
    }
    case kQuarterStart: {
      const cal::year_month ym = cal::year_month{ymd.year(), ymd.month()} + cal::months{3 * k};
      const unsigned m = (static_cast<unsigned>(ym.month()) - 1) / 3 * 3 + 1;
      return cal::sys_days{ym.year() / cal::month{m} / cal::day{1}};
    }
    case kYearStart:
      return cal::sys_days{(ymd.year() + cal::years{k}) / cal::January / cal::day{1}};
    case kWeekStart:
      return d + cal::days{7 * k};
    default:
      return d + cal::days{k};
  }
}

Result<std::shared_ptr<Array>> resample_labels_calendar(const std::shared_ptr<Array>& ax, int type, int multiplier,
                                                        bool closed_right, bool label_right) {
  auto ts = std::dynamic_pointer_cast<arrow::TimestampArray>(ax);
  if (!ts) return Status::Invalid("axis must be a TimestampArray but got array of type ", ax->type()->ToString());
  if (ts->null_count() > 0) return Status::NotImplemented("oracle: null timestamps");
  arrow::TimestampBuilder out(ts->type(), arrow::default_memory_pool());
  if (ts->length() == 0) return out.Finish();
  ARROW_ASSIGN_OR_RAISE(Datum mm, ac::MinMax(ts));
  const auto& st = mm.scalar_as<arrow::StructScalar>();
  const int64_t mn = static_cast<const arrow::TimestampScalar&>(*st.value[0]).value;
  const int64_t mx = static_cast<const arrow::TimestampScalar&>(*st.value[1]).value;
  if (!closed_right) return Status::NotImplemented("closed_left is not currently supported by DateOffset");
  if (multiplier < 1) return Status::Invalid("FREQ must be >= 1");
  switch (type) {   // switchFunction, core.cpp:235-265
    case kMonthEnd: return Status::NotImplemented("MonthEnd not supported use arrow month().groupby()");
    case kQuarterEnd: return Status::NotImplemented("QuarterEnd not supported use arrow quarter().groupby()");
    case kWeekEnd: return Status::NotImplemented("WeekEnd not supported use arrow weeks().groupby()");
    case kYearEnd: return Status::NotImplemented("YearEnd not supported use arrow year().groupby()");
    default: break;
  }
  const cal::sys_days first{cal::days{floor_to_day(mn) / kNsPerDay}}, last{cal::days{floor_to_day(mx) / kNsPerDay}};
  const cal::sys_days start = offset_add(first, type, -multiplier), stop = offset_add(last, type, multiplier);
  if (start >= stop) return Status::Invalid("start date has to be less than end date");
  if (type == kQuarterStart && static_cast<unsigned>(cal::year_month_day{start}.month()) / 3 != 0)
    return Status::Invalid("A quarter freq requires month is on a quarter, +/- with DateOffset");
  std::vector<int64_t> binner;
  {
    cal::sys_days it = start;
    cal::year_month ym{cal::year_month_day{start}.year(), cal::year_month_day{start}.month()};
    while (it <= stop) {
      binner.push_back(static_cast<int64_t>(it.time_since_epoch().count()) * kNsPerDay);
      switch (type) {   // day_iterator / week_iterator / month_iterator (x3 for quarters) / year_iterator
        case kDay: it += cal::days{multiplier}; break;
        case kWeekStart: it += cal::days{7 * multiplier}; break;
        case kMonthStart: ym += cal::months{multiplier}; it = cal::sys_days{ym / cal::day{1}}; break;
        case kQuarterStart: ym += cal::months{3 * multiplier}; it = cal::sys_days{ym / cal::day{1}}; break;
        default: ym += cal::years{multiplier}; it = cal::sys_days{ym / cal::day{1}}; break;
      }
    }
  }
  std::vector<int64_t> edges = binner;
  if (!(type == kDay && multiplier == 1)) {   // adjustBinEdges: + 1 day - 1 ns; drop the last edge when the one before covers max
    for (auto& e : edges) e += kNsPerDay - 1;
    if (edges[edges.size() - 2] > mx) { edges.pop_back(); binner.pop_back(); }
  }
  std::vector<int64_t> bins;
  ARROW_RETURN_NOT_OK(generate_bins(ts->raw_values(), ts->length(), edges, true, &bins));
  size_t label_begin = label_right ? 1 : 0;
  size_t n_labels = binner.size() - label_begin;
  if (bins.size() < n_labels) n_labels = bins.size();
  if (bins.back() < static_cast<int64_t>(n_labels)) return Status::NotImplemented("upSampling is not implemented.");
  if (bins.size() != n_labels) return Status::Invalid("Processing Group Info requires bins.size() == labels->length()");
  std::vector<int64_t> labels(bins.back());
  int64_t prev = 0;
  for (size_t b = 0; b < bins.size(); ++b) {
    for (int64_t i = prev; i < bins[b]; ++i) labels[i] = binner[label_begin + b];
    prev = bins[b];
  }
  ARROW_RETURN_NOT_OK(out.AppendValues(labels));
  return out.Finish();
}

// dataframe.cpp:1265-1290 for fixed-width units (the month/week/year/quarter "subtract one
// day" branch at :1279-1287 is restated too since it is one Arrow call chain).
Result<std::shared_ptr<Array>> downsample_labels(const std::shared_ptr<Array>& index, int multiple,
                                                 char unit, bool closed_label_right,
                                                 bool week_starts_monday, bool start_epoch) {
  ac::CalendarUnit cu;
  switch (unit) {  // core.cpp getCalendarUnit
    case 'N': cu = ac::CalendarUnit::NANOSECOND; break;
    case 'U': cu = ac::CalendarUnit::MICROSECOND; break;
    case 'L': cu = ac::CalendarUnit::MILLISECOND; break;
    case 'S': cu = ac::CalendarUnit::SECOND; break;
    case 'T': cu = ac::CalendarUnit::MINUTE; break;
    case 'H': cu = ac::CalendarUnit::HOUR; break;
    case 'D': cu = ac::CalendarUnit::DAY; break;
    case 'W': cu = ac::CalendarUnit::WEEK; break;
    case 'M': cu = ac::CalendarUnit::MONTH; break;
    case 'Q': cu = ac::CalendarUnit::QUARTER; break;
    case 'Y': cu = ac::CalendarUnit::YEAR; break;
    default: return Status::Invalid("unknown calendar unit");
  }
  ac::RoundTemporalOptions opt(multiple, cu, week_starts_monday, false, start_epoch);
  ARROW_ASSIGN_OR_RAISE(Datum binned, closed_label_right ? ac::CeilTemporal(index, opt)
                                                         : ac::FloorTemporal(index, opt));
  auto arr = binned.make_array();
  if (unit == 'M' || unit == 'W' || unit == 'Y' || unit == 'Q') {
    ARROW_ASSIGN_OR_RAISE(auto one_day, arrow::MakeScalar(arrow::date32(), 1L));
    ARROW_ASSIGN_OR_RAISE(Datum d, ac::Subtract(arr, one_day));
    ARROW_ASSIGN_OR_RAISE(d, ac::Cast(d, arrow::int64()));
    ARROW_ASSIGN_OR_RAISE(d, ac::Cast(d, arrow::timestamp(arrow::TimeUnit::NANO)));
    arr = d.make_array();
  }
  return arr;
}

Status export_array(const std::shared_ptr<Array>& a, ArrowArray* out, ArrowSchema* out_schema) {
  return arrow::ExportArray(*a, out, out_schema);
}

}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// batch/schema: a RecordBatch exported through the C Data Interface (consumed).
// index/index_schema: optional index array (consumed when non-null).
void* orc_groupby_create(ArrowArray* batch, ArrowSchema* schema, ArrowArray* index,
                         ArrowSchema* index_schema, const char** key_names, int n_keys,
                         int materialize) {
  ensure_compute_initialized();
  auto g = std::make_unique<Groups>();
  auto rb = arrow::ImportRecordBatch(batch, schema);
  if (!rb.ok()) { fail(rb.status()); return nullptr; }
  g->frame = *rb;
  if (index && index_schema) {
    auto ix = arrow::ImportArray(index, index_schema);
    if (!ix.ok()) { fail(ix.status()); return nullptr; }
    g->index = *ix;
  }
  for (int i = 0; i < n_keys; ++i) g->key_names.emplace_back(key_names[i]);
  auto st = g->build(materialize != 0);
  if (!st.ok()) { fail(st); return nullptr; }
  return g.release();
}

void orc_groupby_destroy(void* h) { delete static_cast<Groups*>(h); }

int64_t orc_groupby_num_groups(void* h) { return static_cast<Groups*>(h)->num_groups(); }

// t[0..3] = consume, groupings+uniques, gathers+map inserts, total (ms)
void orc_groupby_timing(void* h, double* t) {
  auto& tm = static_cast<Groups*>(h)->timing;
  t[0] = tm.consume_ms; t[1] = tm.groupings_ms; t[2] = tm.gather_ms; t[3] = tm.total_ms;
}

int orc_groupby_unique(void* h, int key_i, ArrowArray* out, ArrowSchema* out_schema) {
  auto* g = static_cast<Groups*>(h);
  if (key_i < 0 || key_i >= static_cast<int>(g->unique_keys.size())) return fail("bad key index");
  auto st = export_array(g->unique_keys[key_i], out, out_schema);
  return st.ok() ? 0 : fail(st);
}

int orc_groupby_row_ids(void* h, ArrowArray* out, ArrowSchema* out_schema) {
  auto* g = static_cast<Groups*>(h);
  auto st = export_array(g->row_ids, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

// func: "sum" "min" "max" "product"  -> GROUPBY_AGG semantics (nulls kept)
//       "mean" "count" "variance" "stddev" "all" "any" -> GROUPBY_NUMERIC_AGG semantics (validity dropped);
//                                       out_valid (optional) receives the scalar validity
//       "first" "last"               -> positional
// out_valid/out_valid_schema may be NULL.
int orc_groupby_agg(void* h, const char* func, const char* column, int nthreads, ArrowArray* out,
                    ArrowSchema* out_schema, ArrowArray* out_valid, ArrowSchema* out_valid_schema) {
  auto* g = static_cast<Groups*>(h);
  const std::string f(func);
  if (nthreads < 1) nthreads = 1;
  Result<std::shared_ptr<Array>> r = Status::NotImplemented("aggregate ", f);
  std::shared_ptr<Array> valid;
  if (f == "sum" || f == "min" || f == "max" || f == "product") r = agg_boxed(*g, f, column, nthreads);
  else if (f == "mean" || f == "count" || f == "variance" || f == "stddev" || f == "all" || f == "any" || f == "count_distinct") r = agg_numeric(*g, f, column, nthreads, out_valid ? &valid : nullptr);
  else if (f == "first") r = agg_position(*g, false, column);
  else if (f == "last") r = agg_position(*g, true, column);
  if (!r.ok()) return fail(r.status());
  auto st = export_array(*r, out, out_schema);
  if (!st.ok()) return fail(st);
  if (out_valid && valid) {
    st = export_array(valid, out_valid, out_valid_schema);
    if (!st.ok()) return fail(st);
  }
  return 0;
}

int orc_groupby_min_max(void* h, const char* column, int nthreads, ArrowArray* out_min,
                        ArrowSchema* out_min_schema, ArrowArray* out_max,
                        ArrowSchema* out_max_schema) {
  auto* g = static_cast<Groups*>(h);
  std::shared_ptr<Array> mn, mx;
  auto st = agg_min_max(*g, column, nthreads < 1 ? 1 : nthreads, &mn, &mx);
  if (!st.ok()) return fail(st);
  st = export_array(mn, out_min, out_min_schema);
  if (!st.ok()) return fail(st);
  st = export_array(mx, out_max, out_max_schema);
  return st.ok() ? 0 : fail(st);
}

// group(value) equivalent for tests: the j-th group's slice of a column (group_by.h:38-50).
int orc_groupby_group_slice(void* h, const char* column, int64_t j, ArrowArray* out,
                            ArrowSchema* out_schema) {
  auto* g = static_cast<Groups*>(h);
  auto c = g->column_index(column);
  if (!c.ok()) return fail(c.status());
  auto st = g->ensure_gathered(*c);
  if (!st.ok()) return fail(st);
  auto s = g->group_slice(*c, j);
  if (!s.ok()) return fail(s.status());
  st = export_array(*s, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

int orc_resample_labels(ArrowArray* index, ArrowSchema* index_schema, int64_t freq_ns,
                        int closed_right, int label_right, int origin, int64_t origin_custom_ns,
                        int64_t offset_ns, ArrowArray* out, ArrowSchema* out_schema) {
  ensure_compute_initialized();
  auto ix = arrow::ImportArray(index, index_schema);
  if (!ix.ok()) return fail(ix.status());
  auto r = resample_labels(*ix, freq_ns, closed_right != 0, label_right != 0, origin,
                           origin_custom_ns, offset_ns);
  if (!r.ok()) return fail(r.status());
  auto st = export_array(*r, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

int orc_resample_labels_calendar(ArrowArray* index, ArrowSchema* index_schema, int offset_type, int multiplier,
                                 int closed_right, int label_right, ArrowArray* out, ArrowSchema* out_schema) {
  ensure_compute_initialized();
  auto ix = arrow::ImportArray(index, index_schema);
  if (!ix.ok()) return fail(ix.status());
  auto r = resample_labels_calendar(*ix, offset_type, multiplier, closed_right != 0, label_right != 0);
  if (!r.ok()) return fail(r.status());
  auto st = export_array(*r, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

// Series::argsort / Series::sort (series.cpp:864-868, 978-992): CallFunction("array_sort_indices", {array},
// ArraySortOptions{order}) and Take(array, indices) — DataFrame::sort_index (dataframe.cpp:1062-1071) applies the same
// indices to every column.  take_sorted = 1 returns Take(values, indices) instead of the indices.
int orc_array_sort(ArrowArray* values, ArrowSchema* schema, int ascending, int take_sorted, ArrowArray* out, ArrowSchema* out_schema) {
  ensure_compute_initialized();
  auto arr = arrow::ImportArray(values, schema);
  if (!arr.ok()) return fail(arr.status());
  ac::ArraySortOptions opt{ascending ? ac::SortOrder::Ascending : ac::SortOrder::Descending};
  auto idx = ac::CallFunction("array_sort_indices", {*arr}, &opt);
  if (!idx.ok()) return fail(idx.status());
  std::shared_ptr<Array> res = idx->make_array();
  if (take_sorted) {
    auto t = ac::Take(*arr, idx->make_array());
    if (!t.ok()) return fail(t.status());
    res = t->make_array();
  }
  auto st = export_array(res, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

int orc_downsample_labels(ArrowArray* index, ArrowSchema* index_schema, int multiple, char unit,
                          int closed_label_right, int week_starts_monday, int start_epoch,
                          ArrowArray* out, ArrowSchema* out_schema) {
  ensure_compute_initialized();
  auto ix = arrow::ImportArray(index, index_schema);
  if (!ix.ok()) return fail(ix.status());
  auto r = downsample_labels(*ix, multiple, unit, closed_label_right != 0,
                             week_starts_monday != 0, start_epoch != 0);
  if (!r.ok()) return fail(r.status());
  auto st = export_array(*r, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

// NDFrame<T>::sum/mean/min/max/count (ndframe.cpp:26-55): CallFunction(name, {array},
// ScalarAggregateOptions{skip_null}) — min_count stays at its default of 1.  Result is a
// length-1 array.  "count" uses CountOptions default (ONLY_VALID), ndframe.cpp:119.
int orc_scalar_agg(ArrowArray* array, ArrowSchema* schema, const char* func, int skip_null,
                   ArrowArray* out, ArrowSchema* out_schema) {
  ensure_compute_initialized();
  auto a = arrow::ImportArray(array, schema);
  if (!a.ok()) return fail(a.status());
  const std::string f(func);
  Result<Datum> d = Status::NotImplemented(f);
  if (f == "count") {
    d = ac::CallFunction("count", {*a});
  } else {
    ac::ScalarAggregateOptions opt(skip_null != 0);
    d = ac::CallFunction(f, {*a}, &opt);
  }
  if (!d.ok()) return fail(d.status());
  auto arr = arrow::MakeArrayFromScalar(*d->scalar(), 1);
  if (!arr.ok()) return fail(arr.status());
  auto st = export_array(*arr, out, out_schema);
  return st.ok() ? 0 : fail(st);
}

}  // extern "C"
