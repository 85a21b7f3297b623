// pd_groupby.h — host-side C++ façade with the reference's class surface for the hot path, bound
// to the CUDA library through the C ABI (include/pa_b200.h) instead of arrow::compute.
//
//   pd::GroupBy     <- /root/reference/src/group_by.h:22-247, dataframe.cpp:1512-1806
//   pd::Resampler   <- /root/reference/src/group_by.h:255-299
//   pd::resample    <- /root/reference/src/resample.h:51-122
//   pd::DataFrame::group_by / resample / downsample   <- dataframe.cpp:1227-1290
//   pd::Series::resample / group_by                   <- series.cpp:351-359 (+ north_star: Series::group_by)
//
// Same names, argument meaning, result shapes and error behaviour (constructors throw
// std::runtime_error, aggregations return arrow::Result) as the reference.  pd::DataFrame and
// pd::Series here are only the thin containers the path needs (a RecordBatch / Array plus an
// index), not the reference's ~300-method wrappers (out of scope).
//
// Group materialisation (group(), MakeSubDataFrame, apply*, orderedGroups; group_by.h:38-83,141-162,
// dataframe.cpp:1354-1510) is built lazily ON THE DEVICE: pa_groupby_groupings (= Grouper::MakeGroupings) and
// pa_groupby_take_grouped (= Grouper::ApplyGroupings), the calls the reference makes after Consume
// (dataframe.cpp:1539-1569,1586-1597); the façade only slices the gathered columns.
// product / variance / stddev run a second device pass (stage2.cuh).  Not carried over (SURVEY.md §8f
// "next"): approximate_median/mode/tdigest; those methods exist and return
// arrow::Status::NotImplemented.
#pragma once
#include <arrow/api.h>

#include <chrono>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

struct pa_groupby;

namespace pd {

using ArrayPtr = std::shared_ptr<arrow::Array>;
using ScalarPtr = std::shared_ptr<arrow::Scalar>;
using time_duration = std::chrono::nanoseconds;   // boost::posix_time::time_duration in the reference

inline time_duration minutes(int64_t n) { return std::chrono::minutes(n); }
inline time_duration seconds(int64_t n) { return std::chrono::seconds(n); }
inline time_duration milliseconds(int64_t n) { return std::chrono::milliseconds(n); }
inline time_duration microseconds(int64_t n) { return std::chrono::microseconds(n); }
inline time_duration nanoseconds(int64_t n) { return std::chrono::nanoseconds(n); }

// core.h:79-91
struct TimeGrouperOrigin {
  enum Type { Epoch, Start, StartDay, End, EndDay, Custom } type{StartDay};
  int64_t custom_ns{0};
};

// core.h:122-159: calendar rule of pd::resample ("2D", "WS", "MS", "QS", "YS"; the END types exist in the reference but
// its date_range rejects them, core.cpp:243-260)
struct DateOffset {
  enum Type { Day, MonthEnd, QuarterStart, QuarterEnd, WeekStart, WeekEnd, MonthStart, YearEnd, YearStart } type{Day};
  int multiplier{1};
  static std::optional<DateOffset> FromString(std::string const& code);   // core.cpp:62-108
};

// core.h:181-194
template <class T>
T ReturnOrThrowOnFailure(arrow::Result<T>&& result) {
  if (result.ok()) return result.MoveValueUnsafe();
  throw std::runtime_error(result.status().ToString());
}
inline void ThrowOnFailure(arrow::Status&& status) {
  if (!status.ok()) throw std::runtime_error(status.ToString());
}

// core.cpp:110-133
std::pair<std::string, int> splitTimeSpan(std::string const& freq);

// scalar.h:62-241 (reading results only)
class Scalar {
 public:
  Scalar() = default;
  explicit Scalar(ScalarPtr s) : scalar(std::move(s)) {}
  template <class T>
  T as() const;
  bool isValid() const { return scalar && scalar->is_valid; }
  ScalarPtr value() const { return scalar; }
  bool operator==(int64_t v) const { return isValid() && as<int64_t>() == v; }
  bool operator==(double v) const { return isValid() && as<double>() == v; }
  ScalarPtr scalar;
};

class GroupBy;
class Resampler;
class DataFrame;

class Series {
 public:
  Series() = default;
  Series(ArrayPtr array, ArrayPtr index, std::string name = "") : m_array(std::move(array)), m_index(std::move(index)), m_name(std::move(name)) {}
  const ArrayPtr& array() const { return m_array; }
  const ArrayPtr& indexArray() const { return m_index; }
  const std::string& name() const { return m_name; }
  std::shared_ptr<arrow::DataType> dtype() const { return m_array->type(); }
  int64_t size() const { return m_array ? m_array->length() : 0; }
  Scalar operator[](int64_t i) const { return Scalar(ReturnOrThrowOnFailure(m_array->GetScalar(i))); }
  template <class T>
  std::vector<T> values() const;
  // NDFrame::sum/mean/min/max/count/first/last/min_max/agg (ndframe.cpp:119,129,160-175,220,237-241; SURVEY §8 row
  // a17) on the GPU: the column is aggregated as ONE group by the fused pass (pa_column_aggregate).  skip_null =
  // false: null as soon as the column holds a null, as ScalarAggregateOptions{skip_nulls = false} makes arrow answer.
  // first / last follow arrow's scalar kernels here (skip_null: first / last VALID value), unlike GroupBy::first.
  Scalar sum(bool skip_null = true) const;
  Scalar mean(bool skip_null = true) const;
  Scalar min(bool skip_null = true) const;
  Scalar max(bool skip_null = true) const;
  Scalar first(bool skip_null = true) const;
  Scalar last(bool skip_null = true) const;
  Scalar product(bool skip_null = true) const;
  Scalar agg(std::string const& name, bool skip_null = true) const;
  std::pair<Scalar, Scalar> min_max(bool skip_null = true) const;
  int64_t count() const;
  Scalar sum_on_device(bool skip_null = true) const;   // (kept for callers of round 1: same as sum())
  // series.cpp:864-868 / 978-992: arrow's array_sort_indices (+ Take) — a stable device argsort here (sort.cuh);
  // sort() throws without an index, like the reference
  Series argsort(bool ascending = true) const;
  Series sort(bool ascending = true) const;
  // series.cpp:351-359
  Resampler resample(std::string const& rule, bool closed_right = false, bool label_right = false,
                     TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(0),
                     std::string const& tz = "") const;
  // north_star: Series::group_by — the analogue of resample.h:109-115 (wrap into a one-column frame)
  GroupBy group_by(const ArrayPtr& key) const;

 private:
  ArrayPtr m_array, m_index;
  std::string m_name;
};

class DataFrame {
 public:
  DataFrame() = default;
  explicit DataFrame(std::shared_ptr<arrow::RecordBatch> rb, ArrayPtr index = nullptr);
  // dataframe.h:87-93
  DataFrame(std::shared_ptr<arrow::Schema> const& schema, int64_t num_rows, arrow::ArrayVector const& arrays, ArrayPtr index = nullptr);
  // (index, pair{name, vector}...) — the form the reference's tests use
  template <class... Pairs>
  DataFrame(ArrayPtr index, Pairs&&... cols) : m_index(std::move(index)) {
    arrow::FieldVector fields;
    arrow::ArrayVector arrays;
    (addColumn(fields, arrays, cols.first, cols.second), ...);
    init(arrow::schema(fields), arrays);
  }
  template <class T>
  explicit DataFrame(std::map<std::string, std::vector<T>> const& cols) {
    arrow::FieldVector fields;
    arrow::ArrayVector arrays;
    for (auto const& [name, v] : cols) addColumn(fields, arrays, name, v);
    init(arrow::schema(fields), arrays);
  }

  const std::shared_ptr<arrow::RecordBatch>& array() const { return m_array; }
  const ArrayPtr& indexArray() const { return m_index; }
  int64_t num_rows() const { return m_array ? m_array->num_rows() : 0; }
  int64_t num_columns() const { return m_array ? m_array->num_columns() : 0; }
  std::vector<std::string> columnNames() const { return m_array->schema()->field_names(); }
  Series operator[](std::string const& name) const;
  Scalar at(int64_t row, int64_t col) const { return Scalar(ReturnOrThrowOnFailure(m_array->column(static_cast<int>(col))->GetScalar(row))); }
  DataFrame setIndex(ArrayPtr const& index) const { return DataFrame(m_array, index); }
  Scalar sum() const;   // ndframe.cpp:220 over all columns concatenated (ndframe.h:329-335)

  // dataframe.cpp:1062-1071: rows reordered by the sorted index (device argsort + scatter-shaped take of every column)
  DataFrame sort_index(bool ascending = true, bool ignore_index = false) const;
  // dataframe.cpp:1188-1208: every column named in `by` is replaced by its own sorted values (the reference sorts the
  // listed columns independently and drops the index)
  DataFrame sort_values(std::vector<std::string> const& by, bool ascending = true) const;
  // dataframe.cpp:757-791 / 646-683: one record batch out of an Arrow IPC stream / a Parquet file; `index` names the
  // column that becomes the index (int64 -> timestamp[ns], as the reference casts it)
  static DataFrame readBinary(std::basic_string_view<uint8_t> const& blob, std::optional<std::string> const& index = std::nullopt);
  static DataFrame readParquet(std::string const& path);
  // Ingest: every fixed-width / boolean / utf8 column and the index are copied to the device ONCE
  // (pa_column_to_device); group_by / resample / sort / the aggregates on the returned frame — and on the Series taken
  // from it — then use the device copies in place instead of uploading a column per call.
  DataFrame to_device() const;
  bool on_device() const { return static_cast<bool>(m_device); }

  // dataframe.cpp:1227-1235
  GroupBy group_by(const std::string& key) const;
  GroupBy group_by(const ArrayPtr& keyArray) const;
  // dataframe.cpp:1254-1262
  Resampler resample(std::string const& rule, bool closed_right = false, bool label_right = false,
                     TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(0),
                     std::string const& tz = "") const;
  // dataframe.cpp:1265-1290: every unit of the reference (N U L S T H D W M Q Y), labels computed on the device
  Resampler downsample(std::string const& rule, bool closed_label_right = true, bool weekStartsMonday = true,
                       bool startEpoch = true) const;

  std::shared_ptr<arrow::RecordBatch> m_array;
  ArrayPtr m_index;
  std::shared_ptr<void> m_device;   // device copies of the columns (to_device()); shared by copies of the frame

 private:
  template <class T>
  static void addColumn(arrow::FieldVector& fields, arrow::ArrayVector& arrays, std::string const& name, std::vector<T> const& v);
  void init(std::shared_ptr<arrow::Schema> schema, arrow::ArrayVector const& arrays);
};

class GroupBy {
 public:
  // group_by.h:24-31 — throws std::runtime_error when the key cannot be grouped
  GroupBy(const std::string& key, DataFrame df);
  GroupBy(GroupBy&&) noexcept;
  GroupBy& operator=(GroupBy&&) noexcept;
  GroupBy(const GroupBy&) = delete;
  ~GroupBy();

  size_t groupSize() const;                                   // group_by.h:33-36
  ArrayPtr unique() const;                                    // group_by.h:52-55
  ScalarPtr GetKeyByIndex(int64_t i) const { return ReturnOrThrowOnFailure(unique()->GetScalar(i)); }

  arrow::Result<DataFrame> mean(std::vector<std::string> const& args);
  arrow::Result<Series> mean(std::string const& arg);
  arrow::Result<DataFrame> count(std::vector<std::string> const& args);
  arrow::Result<Series> count(std::string const& arg);
  arrow::Result<DataFrame> first(std::vector<std::string> const& args);
  arrow::Result<Series> first(std::string const& arg);
  arrow::Result<DataFrame> last(std::vector<std::string> const& args);
  arrow::Result<Series> last(std::string const& arg);
  arrow::Result<DataFrame> max(std::vector<std::string> const& args);
  arrow::Result<Series> max(std::string const& arg);
  arrow::Result<DataFrame> min(std::vector<std::string> const& args);
  arrow::Result<Series> min(std::string const& arg);
  arrow::Result<DataFrame> min_max(std::vector<std::string> const& args);
  arrow::Result<DataFrame> min_max(std::string const& arg);
  arrow::Result<DataFrame> sum(std::vector<std::string> const& args);
  arrow::Result<Series> sum(std::string const& arg);
  arrow::Result<DataFrame> product(std::vector<std::string> const& args);
  arrow::Result<Series> product(std::string const& arg);
  arrow::Result<DataFrame> variance(std::vector<std::string> const& args);
  arrow::Result<Series> variance(std::string const& arg);
  arrow::Result<DataFrame> stddev(std::vector<std::string> const& args);
  arrow::Result<Series> stddev(std::string const& arg);
  arrow::Result<DataFrame> all(std::vector<std::string> const& args);      // boolean columns
  arrow::Result<Series> all(std::string const& arg);
  arrow::Result<DataFrame> any(std::vector<std::string> const& args);
  arrow::Result<Series> any(std::string const& arg);
  arrow::Result<DataFrame> count_distinct(std::vector<std::string> const& args);
  arrow::Result<Series> count_distinct(std::string const& arg);

  // declared by the reference, outside this path's CUDA scope
#define PD_NOT_ON_GPU(name)                                                                                   \
  arrow::Result<DataFrame> name(std::vector<std::string> const&) { return arrow::Status::NotImplemented(#name " is not part of the B200 group-by path"); } \
  arrow::Result<Series> name(std::string const&) { return arrow::Status::NotImplemented(#name " is not part of the B200 group-by path"); }
  PD_NOT_ON_GPU(approximate_median)
  PD_NOT_ON_GPU(mode) PD_NOT_ON_GPU(tdigest)
#undef PD_NOT_ON_GPU
  // ---- materialised groups (group_by.h:38-83, 141-162; dataframe.cpp:1430-1510) ----
  arrow::ArrayVector group(ScalarPtr const& key) const;                       // every column's rows of one group
  template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, ScalarPtr>>>
  arrow::ArrayVector group(T&& value) const {
    ScalarPtr key = arrow::MakeScalar(std::forward<T>(value));
    return group(key);
  }
  DataFrame MakeSubDataFrame(int64_t groupIndex, std::shared_ptr<arrow::Schema> const& schema) const;
  DataFrame MakeSubDataFrame(ScalarPtr const& key, std::shared_ptr<arrow::Schema> const& schema) const;
  arrow::Result<Series> apply(std::function<ScalarPtr(DataFrame const&)> fn);
  arrow::Result<Series> apply(std::function<ArrayPtr(DataFrame const&)> fn);
  arrow::Result<DataFrame> apply(std::function<ScalarPtr(Series const&)> fn);
  // group_by.h:77, dataframe.cpp:1411-1428: fn on every group's sub-frame, results concatenated row-wise (pd::concat along
  // the index; the results must share one schema)
  arrow::Result<DataFrame> apply_chunk(std::function<DataFrame(DataFrame const&)> fn);
  arrow::Result<Series> apply_async(std::function<ScalarPtr(DataFrame const&)> fn) { return apply(std::move(fn)); }
  arrow::Result<DataFrame> apply_async(std::function<ScalarPtr(Series const&)> fn) { return apply(std::move(fn)); }
  template <class IndexType>
  std::vector<std::pair<IndexType, DataFrame>> orderedGroups() const {
    const int64_t n = static_cast<int64_t>(groupSize());
    auto schema = df.m_array->schema();
    std::vector<std::pair<IndexType, DataFrame>> result;
    result.reserve(n);
    for (int64_t g = 0; g < n; ++g) result.emplace_back(Scalar(GetKeyByIndex(g)).template as<IndexType>(), MakeSubDataFrame(g, schema));
    return result;
  }

 protected:
  GroupBy() = default;
  const DataFrame& getDF() const { return df; }
  // one fused pass for `column`, returns the arrays of the aggregates in `mask` (ascending bit order)
  arrow::Result<arrow::ArrayVector> aggregate(std::string const& column, uint32_t mask, bool drop_validity);
  arrow::Result<DataFrame> frameOf(std::vector<std::string> const& args, uint32_t bit, bool drop_validity, bool with_index);
  arrow::Result<Series> seriesOf(std::string const& arg, uint32_t bit, bool drop_validity, bool with_index);

  DataFrame df;
  pa_groupby* handle = nullptr;
  ArrayPtr key_array;                          // kept alive: the C ABI borrows the key buffers
  std::shared_ptr<arrow::Array> key_dictionary;  // for dictionary / utf8 keys
  mutable ArrayPtr uniqueKeys;
  // lazily materialised groups, by group index (first-appearance order)
  void materialize() const;
  int64_t indexOfKey(ScalarPtr const& key) const;
  mutable bool materialized = false;
  mutable std::vector<arrow::ArrayVector> groups;     // groups[g][column]
  mutable arrow::ArrayVector indexGroups;             // indexGroups[g]
  friend class Resampler;
};

// group_by.h:255-299
class Resampler : protected GroupBy {
 public:
  explicit Resampler(DataFrame const& _df);    // hash group-by on the index ("__resampler_idx__")
  // time-bucket specialisation: sorted index + fixed width (pd::resample)
  Resampler(DataFrame const& _df, int64_t freq_ns, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
            int64_t offset_ns);
  // calendar rule (makeGroupInfo's DateOffset branch, resample.cpp:248-267): edges / labels made by the host side of the
  // C ABI, rows reduced by the same sorted-run kernel
  Resampler(DataFrame const& _df, DateOffset const& rule, bool closed_right, bool label_right);
  // DataFrame::downsample (dataframe.cpp:1265-1290): groups on Floor/CeilTemporal(index) labels made on the device
  struct DownsampleRule { int multiple; char unit; bool closed_label_right, week_starts_monday, calendar_based_origin; };
  Resampler(DataFrame const& _df, DownsampleRule const& rule);
  ArrayPtr index() const { return this->unique(); }
  const DataFrame& data() const { return this->getDF(); }
#define PD_RESAMPLE_FN(name) \
  arrow::Result<DataFrame> name() { return frameOfAll(#name); }
  PD_RESAMPLE_FN(mean) PD_RESAMPLE_FN(count) PD_RESAMPLE_FN(max) PD_RESAMPLE_FN(min) PD_RESAMPLE_FN(sum)
  PD_RESAMPLE_FN(first) PD_RESAMPLE_FN(last) PD_RESAMPLE_FN(product) PD_RESAMPLE_FN(variance) PD_RESAMPLE_FN(stddev)
#undef PD_RESAMPLE_FN
  arrow::Result<DataFrame> min_max() { return GroupBy::min_max(getDF().columnNames()); }
  using GroupBy::apply;
  using GroupBy::groupSize;

 private:
  arrow::Result<DataFrame> frameOfAll(std::string const& name);
};

// resample.h:51-122
Resampler resample(DataFrame const& df, std::string const& rule, bool closed_right = false, bool label_right = false,
                   TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(0), std::string const& tz = "");
Resampler resample(DataFrame const& df, DateOffset const& rule, bool closed_right = false, bool label_right = false,
                   TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(), std::string const& tz = "");
Resampler resample(Series const& s, DateOffset const& rule, bool closed_right = false, bool label_right = false,
                   TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(), std::string const& tz = "");
Resampler resample(DataFrame const& df, time_duration const& rule, bool closed_right = false, bool label_right = false,
                   TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(0), std::string const& tz = "");
Resampler resample(Series const& s, std::string const& rule, bool closed_right = false, bool label_right = false,
                   TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(0), std::string const& tz = "");
Resampler resample(Series const& s, time_duration const& rule, bool closed_right = false, bool label_right = false,
                   TimeGrouperOrigin const& origin = {}, time_duration const& offset = time_duration(0), std::string const& tz = "");

// helpers the reference's tests use (core.cpp:174-425)
ArrayPtr range(int64_t start, int64_t end);                                   // int64 array [start, end)
ArrayPtr date_range(int64_t start_ns, int periods, time_duration freq = std::chrono::minutes(1));
int64_t ns_from_ymd(int y, int m, int d);

// ---------------------------------------------------------------------------------------------
template <class T>
void DataFrame::addColumn(arrow::FieldVector& fields, arrow::ArrayVector& arrays, std::string const& name, std::vector<T> const& v) {
  typename arrow::CTypeTraits<T>::BuilderType b;
  ThrowOnFailure(b.AppendValues(v));
  ArrayPtr a;
  ThrowOnFailure(b.Finish(&a));
  fields.push_back(arrow::field(name, a->type()));
  arrays.push_back(a);
}

template <class T>
T Scalar::as() const {
  if (!scalar) throw std::runtime_error("empty scalar");
  if constexpr (std::is_same_v<T, std::string>) {
    return scalar->ToString();
  } else {
    auto cast = ReturnOrThrowOnFailure(scalar->CastTo(arrow::CTypeTraits<T>::type_singleton()));
    return static_cast<const typename arrow::CTypeTraits<T>::ScalarType&>(*cast).value;
  }
}

template <class T>
std::vector<T> Series::values() const {
  std::vector<T> out(m_array->length());
  for (int64_t i = 0; i < m_array->length(); ++i) out[i] = Scalar(ReturnOrThrowOnFailure(m_array->GetScalar(i))).as<T>();
  return out;
}

}  // namespace pd
