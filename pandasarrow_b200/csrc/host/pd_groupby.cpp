// Host-side façade over the C ABI — see pd_groupby.h.  No computation on the path happens here:
// columns are exported through the Arrow C Data Interface, results imported back.
#include "pd_groupby.h"

#include <arrow/c/bridge.h>
#include <arrow/compute/api.h>
#include <arrow/compute/row/grouper.h>
#include <arrow/io/api.h>
#include <arrow/ipc/api.h>
#include <arrow/table.h>
#include <parquet/arrow/reader.h>

#include <algorithm>
#include <cctype>
#include <iostream>
#include <limits>
#include <map>
#include <mutex>
#include <tuple>
#include <stdexcept>

#include "../../../include/pa_b200.h"

namespace pd {

namespace {

// arrow::compute's function registry (Take for dictionary keys / string group slices, and whatever a caller's apply
// functors use — "add", "multiply" ...) has to be initialised once per process in Arrow >= 21.
const bool g_compute_initialised = [] { return arrow::compute::Initialize().ok(); }();

[[noreturn]] void throw_pa(const char* what) { throw std::runtime_error(std::string(what) + ": " + pa_last_error()); }

arrow::Status pa_status(const char* what) { return arrow::Status::Invalid(what, ": ", pa_last_error()); }

// Device copies made by DataFrame::to_device(), found again by (data buffer address, offset, length) of the host
// array: a Series or a key column taken from an ingested frame shares its buffers, so every export below sees the copy.
struct DeviceColumn {
  ArrowDeviceArray dev{};
  std::shared_ptr<arrow::ArrayData> host;     // keeps the host buffers (and therefore the key) alive
  ~DeviceColumn() { if (dev.array.release) dev.array.release(&dev.array); }
};
using DeviceKey = std::tuple<const void*, int64_t, int64_t>;
std::mutex g_device_mutex;
std::map<DeviceKey, std::weak_ptr<DeviceColumn>> g_device_registry;

DeviceKey device_key(const arrow::ArrayData& d) {
  const void* p = d.buffers.size() > 1 && d.buffers[1] ? static_cast<const void*>(d.buffers[1]->data()) : nullptr;
  return {p, d.offset, d.length};
}
std::shared_ptr<DeviceColumn> find_device_copy(const arrow::Array& a) {
  std::lock_guard<std::mutex> lock(g_device_mutex);
  auto it = g_device_registry.find(device_key(*a.data()));
  if (it == g_device_registry.end()) return nullptr;
  auto sp = it->second.lock();
  if (!sp) g_device_registry.erase(it);
  return sp;
}
struct DeviceColumns {
  std::vector<std::shared_ptr<DeviceColumn>> cols;
  ~DeviceColumns() {
    std::lock_guard<std::mutex> lock(g_device_mutex);
    for (auto& c : cols) g_device_registry.erase(device_key(*c->host));
  }
};

// RAII export of one Arrow array as ArrowDeviceArray + ArrowSchema: the device copy when the array was ingested
// (to_device()), else the host buffers (copied by the library)
struct Exported {
  ArrowDeviceArray dev{};
  ArrowSchema schema{};
  std::shared_ptr<DeviceColumn> resident;
  explicit Exported(const arrow::Array& a) {
    resident = find_device_copy(a);
    if (resident) {
      ThrowOnFailure(arrow::ExportType(*a.type(), &schema));
      dev = resident->dev;
      dev.array.release = nullptr;              // borrowed: the registry entry owns the device buffers
      return;
    }
    ThrowOnFailure(arrow::ExportArray(a, &dev.array, &schema));
    dev.device_id = -1;
    dev.device_type = ARROW_DEVICE_CPU;
  }
  ~Exported() {
    if (dev.array.release) dev.array.release(&dev.array);
    if (schema.release) schema.release(&schema);
  }
  Exported(const Exported&) = delete;
};

arrow::Result<ArrayPtr> import_result(ArrowArray* a, ArrowSchema* s) { return arrow::ImportArray(a, s); }

ArrayPtr strip_validity(const ArrayPtr& a) {
  // GROUPBY_NUMERIC_AGG (pd_core_macros.h:32,67) copies `.value` of every per-group scalar into a plain
  // std::vector<T>, so nulls (all-null groups) come out as valid zeros.
  if (a->null_count() == 0) return a;
  auto d = a->data()->Copy();
  d->buffers[0] = nullptr;
  d->null_count = 0;
  return arrow::MakeArray(d);
}

constexpr uint32_t kSum = PA_AGG_SUM, kMean = PA_AGG_MEAN, kCount = PA_AGG_COUNT, kMin = PA_AGG_MIN, kMax = PA_AGG_MAX,
                   kFirst = PA_AGG_FIRST, kLast = PA_AGG_LAST;

int origin_code(TimeGrouperOrigin::Type t) {
  switch (t) {
    case TimeGrouperOrigin::Epoch: return 0;
    case TimeGrouperOrigin::Start: return 1;
    case TimeGrouperOrigin::StartDay: return 2;
    case TimeGrouperOrigin::End: return 3;
    case TimeGrouperOrigin::EndDay: return 4;
    default: return 5;
  }
}

}  // namespace

std::pair<std::string, int> splitTimeSpan(std::string const& freq) {
  auto it = std::find_if(freq.begin(), freq.end(), [](unsigned char c) { return std::isalpha(c); });
  std::string unit(it, freq.end());
  int value = 1;
  if (it != freq.begin()) value = std::stoi(std::string(freq.begin(), it));
  else if (std::any_of(freq.begin(), freq.end(), [](unsigned char c) { return std::isdigit(c); }))
    throw std::runtime_error("Invalid time offset " + freq);
  return {unit, value};
}

// ------------------------------ containers ------------------------------
DataFrame::DataFrame(std::shared_ptr<arrow::RecordBatch> rb, ArrayPtr index) : m_array(std::move(rb)), m_index(std::move(index)) {
  if (m_array && !m_index) m_index = range(0, m_array->num_rows());
}

DataFrame::DataFrame(std::shared_ptr<arrow::Schema> const& schema, int64_t num_rows, arrow::ArrayVector const& arrays, ArrayPtr index)
    : m_array(arrow::RecordBatch::Make(schema, num_rows, arrays)), m_index(std::move(index)) {
  if (!m_index) m_index = range(0, num_rows);
}

void DataFrame::init(std::shared_ptr<arrow::Schema> schema, arrow::ArrayVector const& arrays) {
  const int64_t n = arrays.empty() ? 0 : arrays[0]->length();
  m_array = arrow::RecordBatch::Make(std::move(schema), n, arrays);
  if (!m_index) m_index = range(0, n);
}

Series DataFrame::operator[](std::string const& name) const {
  auto col = m_array->GetColumnByName(name);
  if (!col) throw std::runtime_error("Invalid column: " + name);
  return Series(col, m_index, name);
}

GroupBy DataFrame::group_by(const std::string& key) const { return GroupBy(key, *this); }

GroupBy DataFrame::group_by(const ArrayPtr& keyArray) const {
  const auto key = "__RESERVED_GROUP_KEY__";
  auto rb = ReturnOrThrowOnFailure(m_array->AddColumn(static_cast<int>(num_columns()), key, keyArray));
  return GroupBy(key, DataFrame(rb, m_index));
}

GroupBy Series::group_by(const ArrayPtr& key) const {
  DataFrame df(arrow::schema({arrow::field(m_name, m_array->type())}), m_array->length(), {m_array}, m_index);
  return df.group_by(key);
}

Resampler DataFrame::resample(std::string const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                              time_duration const& offset, std::string const& tz) const {
  return pd::resample(*this, rule, closed_right, label_right, origin, offset, tz);
}

Resampler Series::resample(std::string const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                           time_duration const& offset, std::string const& tz) const {
  return pd::resample(*this, rule, closed_right, label_right, origin, offset, tz);
}

Resampler DataFrame::downsample(std::string const& rule, bool closed_label_right, bool weekStartsMonday, bool startEpoch) const {
  // dataframe.cpp:1265-1290: per-row label = Floor/CeilTemporal(index, RoundTemporalOptions(value, unit,
  // weekStartsMonday, false, startEpoch)), minus one day for W / M / Q / Y — computed on the device
  // (pa_downsample_create, csrc/temporal.cuh), then the GPU hash group-by on the labels.
  const auto [unit, value] = splitTimeSpan(rule);
  const char u = unit.empty() ? ' ' : unit[0];
  if (std::string("NULSTHDWMQY").find(u) == std::string::npos) throw std::runtime_error("Invalid time offset " + rule);
  return Resampler(*this, Resampler::DownsampleRule{value, u, closed_label_right, weekStartsMonday, startEpoch});
}

// ------------------------------ GroupBy ------------------------------
GroupBy::GroupBy(const std::string& key, DataFrame frame) : df(std::move(frame)) {
  if (!df.m_array) return;   // dataframe.cpp:1574-1576
  if (key == "__resampler_idx__") key_array = df.indexArray();
  else key_array = df[key].array();   // throws std::runtime_error for an unknown column
  if (!key_array) throw std::runtime_error("frame has no index");
  ArrayPtr to_export = key_array;
  // utf8 / large_utf8 keys go to the device as they are (format "u" / "U": offsets + bytes, hashed and verified there,
  // csrc/strkeys.cuh); unique() comes back as strings.
  if (to_export->type_id() == arrow::Type::DICTIONARY) {
    auto d = std::static_pointer_cast<arrow::DictionaryArray>(to_export);
    key_dictionary = d->dictionary();
    key_array = to_export;
  }
  Exported k(*to_export);
  pa_options opt;
  pa_options_init(&opt);
  if (pa_groupby_create(&k.dev, &k.schema, 1, &opt, &handle) != PA_OK) throw_pa("GroupBy");
}

GroupBy::GroupBy(GroupBy&& o) noexcept { *this = std::move(o); }
GroupBy& GroupBy::operator=(GroupBy&& o) noexcept {
  if (this != &o) {
    if (handle) pa_groupby_destroy(handle);
    df = std::move(o.df); handle = o.handle; o.handle = nullptr;
    key_array = std::move(o.key_array); key_dictionary = std::move(o.key_dictionary); uniqueKeys = std::move(o.uniqueKeys);
  }
  return *this;
}
GroupBy::~GroupBy() {
  if (handle) pa_groupby_destroy(handle);
}

size_t GroupBy::groupSize() const {
  if (!handle) return 0;
  int64_t n = 0;
  if (pa_groupby_num_groups(handle, &n) != PA_OK) throw_pa("groupSize");
  return static_cast<size_t>(n);
}

ArrayPtr GroupBy::unique() const {
  if (uniqueKeys || !handle) return uniqueKeys;
  ArrowArray a;
  ArrowSchema s;
  if (pa_groupby_unique(handle, 0, &a, &s) != PA_OK) throw_pa("unique");
  auto arr = ReturnOrThrowOnFailure(import_result(&a, &s));
  if (key_dictionary) {
    // back to the key's own type: take(dictionary, indices) — G elements
    static const bool init = [] { return arrow::compute::Initialize().ok(); }();
    (void)init;
    arr = ReturnOrThrowOnFailure(arrow::compute::Take(key_dictionary, arr)).make_array();
  }
  uniqueKeys = arr;
  return uniqueKeys;
}

// ------------------------------ materialised groups ------------------------------
// The reference materialises every group of every column in its constructor (dataframe.cpp:1571-1600);
// here it happens on first use only.  Group ids come from the GPU, the regrouping is Arrow's.
// Fixed-width numeric / temporal and boolean columns are regrouped on the device (pa_groupby_take_grouped); strings
// and dictionaries with arrow's Take on the row order the device produced.
static bool device_takeable(const arrow::DataType& t) {
  switch (t.id()) {
    case arrow::Type::INT8: case arrow::Type::INT16: case arrow::Type::INT32: case arrow::Type::INT64:
    case arrow::Type::UINT8: case arrow::Type::UINT16: case arrow::Type::UINT32: case arrow::Type::UINT64:
    case arrow::Type::FLOAT: case arrow::Type::DOUBLE: case arrow::Type::TIMESTAMP: case arrow::Type::DATE32:
    case arrow::Type::DATE64: case arrow::Type::TIME32: case arrow::Type::TIME64: case arrow::Type::DURATION:
    case arrow::Type::BOOL:
      return true;
    default:
      return false;
  }
}

// The reference's makeGroups tail (dataframe.cpp:1586-1597): MakeGroupings + ApplyGroupings of the index and of
// every column, here as pa_groupby_groupings (device: row ids -> stable sort) + pa_groupby_take_grouped.
void GroupBy::materialize() const {
  if (materialized) return;
  if (!handle) throw std::runtime_error("GroupBy is empty");
  const int64_t G = static_cast<int64_t>(groupSize());
  if (G > std::numeric_limits<int32_t>::max()) throw std::runtime_error("too many groups to materialise");
  ArrowArray oa, ra;
  ArrowSchema os, rs;
  if (pa_groupby_groupings(handle, &oa, &os, &ra, &rs) != PA_OK) throw_pa("pa_groupby_groupings");
  auto offsets = std::static_pointer_cast<arrow::Int32Array>(ReturnOrThrowOnFailure(arrow::ImportArray(&oa, &os)));
  auto rows = ReturnOrThrowOnFailure(arrow::ImportArray(&ra, &rs));
  auto regroup = [&](const ArrayPtr& col, std::vector<ArrayPtr>* out) {
    ArrayPtr gathered;
    if (device_takeable(*col->type())) {
      Exported v(*col);
      ArrowArray a;
      ArrowSchema sc;
      if (pa_groupby_take_grouped(handle, &v.dev, &v.schema, &a, &sc) != PA_OK) throw_pa("pa_groupby_take_grouped");
      gathered = ReturnOrThrowOnFailure(import_result(&a, &sc));
      if (!gathered->type()->Equals(col->type())) gathered = ReturnOrThrowOnFailure(gathered->View(col->type()));
    } else {
      gathered = ReturnOrThrowOnFailure(arrow::compute::Take(*col, *rows));
    }
    out->resize(G);
    for (int64_t g = 0; g < G; ++g) (*out)[g] = gathered->Slice(offsets->Value(g), offsets->Value(g + 1) - offsets->Value(g));
  };
  regroup(df.indexArray(), &indexGroups);
  groups.assign(G, arrow::ArrayVector{});
  for (int c = 0; c < df.m_array->num_columns(); ++c) {
    std::vector<ArrayPtr> per_group;
    regroup(df.m_array->column(c), &per_group);
    for (int64_t g = 0; g < G; ++g) groups[g].push_back(per_group[g]);
  }
  materialized = true;
}

int64_t GroupBy::indexOfKey(ScalarPtr const& key) const {
  auto keys = unique();
  ScalarPtr k = key;
  if (!k->type->Equals(keys->type())) k = ReturnOrThrowOnFailure(k->CastTo(keys->type()));   // HashScalar: Equals(CastTo(a.type))
  for (int64_t g = 0; g < keys->length(); ++g) {
    auto s = ReturnOrThrowOnFailure(keys->GetScalar(g));
    if (s->Equals(*k)) return g;
  }
  return -1;
}

arrow::ArrayVector GroupBy::group(ScalarPtr const& key) const {
  materialize();
  const int64_t g = indexOfKey(key);
  if (g < 0) {
    std::cout << key->ToString() << " is an invalid key\n";     // group_by.h:45-48
    throw std::out_of_range("invalid group key");
  }
  return groups[g];
}

DataFrame GroupBy::MakeSubDataFrame(int64_t groupIndex, std::shared_ptr<arrow::Schema> const& schema) const {
  materialize();
  if (groupIndex < 0 || groupIndex >= static_cast<int64_t>(groups.size())) throw std::out_of_range("group index");
  return DataFrame(schema, indexGroups[groupIndex]->length(), groups[groupIndex], indexGroups[groupIndex]);
}

DataFrame GroupBy::MakeSubDataFrame(ScalarPtr const& key, std::shared_ptr<arrow::Schema> const& schema) const {
  materialize();
  const int64_t g = indexOfKey(key);
  if (g < 0) throw std::out_of_range("invalid group key");
  return MakeSubDataFrame(g, schema);
}

arrow::Result<DataFrame> GroupBy::apply_chunk(std::function<DataFrame(DataFrame const&)> fn) {
  const std::shared_ptr<arrow::Schema> schema = df.m_array->schema();
  const int64_t G = static_cast<int64_t>(groupSize());
  if (G == 0) return arrow::Status::Invalid("no groups");
  std::vector<DataFrame> parts;
  parts.reserve(G);
  for (int64_t g = 0; g < G; ++g) parts.push_back(fn(MakeSubDataFrame(g, schema)));
  // pd::concat(frames, AxisType::Index) for frames of one schema: column-wise arrow::Concatenate, index included
  const auto out_schema = parts[0].array()->schema();
  arrow::ArrayVector index_parts;
  std::vector<arrow::ArrayVector> col_parts(out_schema->num_fields());
  int64_t rows = 0;
  for (auto const& p : parts) {
    if (!p.array()->schema()->Equals(*out_schema, false)) return arrow::Status::Invalid("apply_chunk: the callback returned frames of different schemas");
    index_parts.push_back(p.indexArray());
    for (int c = 0; c < out_schema->num_fields(); ++c) col_parts[c].push_back(p.array()->column(c));
    rows += p.num_rows();
  }
  arrow::ArrayVector cols;
  for (auto const& cp : col_parts) { ARROW_ASSIGN_OR_RAISE(auto col, arrow::Concatenate(cp)); cols.push_back(col); }
  ARROW_ASSIGN_OR_RAISE(auto index, arrow::Concatenate(index_parts));
  return DataFrame(out_schema, rows, cols, index);
}

static arrow::Result<ArrayPtr> build_array(arrow::ScalarVector const& scalars) {   // group_by.h:191-217
  if (scalars.empty()) return arrow::Status::Invalid("no groups");
  ARROW_ASSIGN_OR_RAISE(auto builder, arrow::MakeBuilder(scalars.back()->type));
  ARROW_RETURN_NOT_OK(builder->AppendScalars(scalars));
  return builder->Finish();
}

arrow::Result<Series> GroupBy::apply(std::function<ScalarPtr(DataFrame const&)> fn) {
  const int64_t G = static_cast<int64_t>(groupSize());
  auto schema = df.m_array->schema();
  arrow::ScalarVector result(G);
  for (int64_t g = 0; g < G; ++g) result[g] = fn(MakeSubDataFrame(g, schema));
  ARROW_ASSIGN_OR_RAISE(auto arr, build_array(result));
  return Series(arr, unique());
}

arrow::Result<Series> GroupBy::apply(std::function<ArrayPtr(DataFrame const&)> fn) {
  const int64_t G = static_cast<int64_t>(groupSize());
  auto schema = df.m_array->schema();
  arrow::ArrayVector result(G);
  for (int64_t g = 0; g < G; ++g) {
    auto sub = MakeSubDataFrame(g, schema);
    result[g] = fn(sub);
    if (result[g]->length() != sub.num_rows())
      throw std::runtime_error("Failed to Merge Apply::Functor due to inconsistent Row Length\n" + std::to_string(result[g]->length()) +
                               " != " + std::to_string(sub.num_rows()));
  }
  ARROW_ASSIGN_OR_RAISE(auto arr, arrow::Concatenate(result));
  return Series(arr, df.indexArray());
}

arrow::Result<DataFrame> GroupBy::apply(std::function<ScalarPtr(Series const&)> fn) {
  materialize();
  const int64_t G = static_cast<int64_t>(groupSize());
  auto schema = df.m_array->schema();
  auto names = schema->field_names();
  arrow::ArrayVector columns;
  arrow::FieldVector fields;
  for (int c = 0; c < schema->num_fields(); ++c) {
    arrow::ScalarVector result(G);
    for (int64_t g = 0; g < G; ++g) result[g] = fn(Series(groups[g][c], indexGroups[g], names[c]));
    ARROW_ASSIGN_OR_RAISE(auto arr, build_array(result));
    fields.push_back(arrow::field(names[c], arr->type()));
    columns.push_back(arr);
  }
  return DataFrame(arrow::schema(fields), G, columns);
}

namespace {
// NDFrame aggregates (ndframe.cpp:119-241): the whole column as ONE group, pa_column_aggregate -> G <= 1 results.
arrow::Result<arrow::ArrayVector> column_aggregate(const ArrayPtr& col, uint32_t mask, bool skip_null = true) {
  if (!col) return arrow::Status::Invalid("empty series");
  pa_options opt;
  pa_options_init(&opt);
  pa_groupby* h = nullptr;
  Exported v(*col);
  if (pa_column_aggregate(&v.dev, &v.schema, mask, skip_null ? 1 : 0, &opt, &h) != PA_OK) return pa_status("pa_column_aggregate");
  std::unique_ptr<pa_groupby, void (*)(pa_groupby*)> guard(h, pa_groupby_destroy);
  arrow::ArrayVector out;
  for (uint32_t bit = 1; bit <= PA_AGG_STDDEV; bit <<= 1) {
    if (!(mask & bit)) continue;
    ArrowArray a;
    ArrowSchema s;
    if (pa_groupby_fetch(h, bit, &a, &s) != PA_OK) return pa_status("fetch");
    ARROW_ASSIGN_OR_RAISE(auto arr, import_result(&a, &s));
    out.push_back(arr);
  }
  return out;
}

// first (only) group of a G <= 1 result; an empty column has no group: null, as arrow's min_count = 1 gives
ScalarPtr only_value(const ArrayPtr& arr) {
  if (arr->length() == 0) return arrow::MakeNullScalar(arr->type());
  return ReturnOrThrowOnFailure(arr->GetScalar(0));
}

Scalar column_scalar(const ArrayPtr& col, uint32_t bit, bool skip_null) {
  auto r = ReturnOrThrowOnFailure(column_aggregate(col, bit, skip_null));
  ScalarPtr s = only_value(r[0]);
  // ScalarAggregateOptions{skip_nulls = false}: any null in the column makes the aggregate null (first / last are
  // positional in that mode and carry their own validity)
  if (!skip_null && col->null_count() > 0 && !(bit & (PA_AGG_FIRST | PA_AGG_LAST))) s = arrow::MakeNullScalar(s->type);
  return Scalar(s);
}

uint32_t agg_bit_of(std::string const& name) {
  if (name == "sum") return PA_AGG_SUM;
  if (name == "mean") return PA_AGG_MEAN;
  if (name == "min") return PA_AGG_MIN;
  if (name == "max") return PA_AGG_MAX;
  if (name == "first") return PA_AGG_FIRST;
  if (name == "last") return PA_AGG_LAST;
  if (name == "product") return PA_AGG_PRODUCT;
  return 0;
}
}  // namespace

Scalar Series::sum(bool skip_null) const { return column_scalar(m_array, PA_AGG_SUM, skip_null); }
Scalar Series::mean(bool skip_null) const { return column_scalar(m_array, PA_AGG_MEAN, skip_null); }
Scalar Series::min(bool skip_null) const { return column_scalar(m_array, PA_AGG_MIN, skip_null); }
Scalar Series::max(bool skip_null) const { return column_scalar(m_array, PA_AGG_MAX, skip_null); }
Scalar Series::first(bool skip_null) const { return column_scalar(m_array, PA_AGG_FIRST, skip_null); }
Scalar Series::last(bool skip_null) const { return column_scalar(m_array, PA_AGG_LAST, skip_null); }
Scalar Series::product(bool skip_null) const { return column_scalar(m_array, PA_AGG_PRODUCT, skip_null); }
Scalar Series::sum_on_device(bool skip_null) const { return sum(skip_null); }

// NDFrame::agg(name, skip_null) (ndframe.cpp:237-241): the aggregates of this path by name
Scalar Series::agg(std::string const& name, bool skip_null) const {
  if (name == "count") return Scalar(arrow::MakeScalar(count()));
  const uint32_t bit = agg_bit_of(name);
  if (!bit) throw std::runtime_error("NotImplemented: Series::agg(\"" + name + "\") is outside the B200 group-by path");
  return column_scalar(m_array, bit, skip_null);
}

std::pair<Scalar, Scalar> Series::min_max(bool skip_null) const {
  auto r = ReturnOrThrowOnFailure(column_aggregate(m_array, PA_AGG_MIN | PA_AGG_MAX));
  ScalarPtr mn = only_value(r[0]), mx = only_value(r[1]);
  if (!skip_null && m_array->null_count() > 0) { mn = arrow::MakeNullScalar(mn->type); mx = arrow::MakeNullScalar(mx->type); }
  return {Scalar(mn), Scalar(mx)};
}

int64_t Series::count() const {
  auto r = ReturnOrThrowOnFailure(column_aggregate(m_array, PA_AGG_COUNT));
  if (r[0]->length() == 0) return 0;
  return std::static_pointer_cast<arrow::Int64Array>(r[0])->Value(0);
}

// ndframe.cpp:220 over all columns (ndframe.h:329-335 concatenates them into one chunked array, which requires
// one common dtype): the device sum of every column, folded in column order like arrow folds the chunks.
Scalar DataFrame::sum() const {
  ScalarPtr total;
  for (int c = 0; c < m_array->num_columns(); ++c) {
    auto r = ReturnOrThrowOnFailure(column_aggregate(m_array->column(c), PA_AGG_SUM));
    ScalarPtr s = only_value(r[0]);
    if (!s->is_valid) continue;
    if (!total) { total = s; continue; }
    if (!total->type->Equals(s->type)) throw std::runtime_error("DataFrame::sum: columns of different types");
    switch (s->type->id()) {
      case arrow::Type::DOUBLE:
        total = arrow::MakeScalar(static_cast<const arrow::DoubleScalar&>(*total).value + static_cast<const arrow::DoubleScalar&>(*s).value);
        break;
      case arrow::Type::INT64:
        total = arrow::MakeScalar(static_cast<int64_t>(static_cast<uint64_t>(static_cast<const arrow::Int64Scalar&>(*total).value) +
                                                       static_cast<uint64_t>(static_cast<const arrow::Int64Scalar&>(*s).value)));
        break;
      default:
        total = arrow::MakeScalar(static_cast<const arrow::UInt64Scalar&>(*total).value + static_cast<const arrow::UInt64Scalar&>(*s).value);
    }
  }
  if (!total) {
    if (m_array->num_columns() == 0) return Scalar(arrow::MakeNullScalar(arrow::float64()));
    auto r = ReturnOrThrowOnFailure(column_aggregate(m_array->column(0), PA_AGG_SUM));
    return Scalar(arrow::MakeNullScalar(r[0]->type()));
  }
  return Scalar(total);
}

arrow::Result<arrow::ArrayVector> GroupBy::aggregate(std::string const& column, uint32_t mask, bool drop_validity) {
  if (!handle) return arrow::Status::Invalid("GroupBy on an empty frame");
  auto col = df.m_array->GetColumnByName(column);
  if (!col) return arrow::Status::KeyError("Invalid column: ", column);
  Exported v(*col);
  if (pa_groupby_aggregate(handle, &v.dev, &v.schema, mask) != PA_OK) return pa_status("aggregate");
  arrow::ArrayVector out;
  for (uint32_t bit = 1; bit <= PA_AGG_COUNT_DISTINCT; bit <<= 1) {
    if (!(mask & bit)) continue;
    ArrowArray a;
    ArrowSchema s;
    if (pa_groupby_fetch(handle, bit, &a, &s) != PA_OK) return pa_status("fetch");
    ARROW_ASSIGN_OR_RAISE(auto arr, import_result(&a, &s));
    out.push_back(drop_validity ? strip_validity(arr) : arr);
  }
  return out;
}

arrow::Result<DataFrame> GroupBy::frameOf(std::vector<std::string> const& args, uint32_t bit, bool drop_validity, bool with_index) {
  arrow::FieldVector fields;
  arrow::ArrayVector arrays;
  for (auto const& arg : args) {
    ARROW_ASSIGN_OR_RAISE(auto r, aggregate(arg, bit, drop_validity));
    // the reference labels the result with the INPUT field (pd_core_macros.h:11,44,85,111) even when the
    // aggregate widened the dtype; the data dtype is authoritative (its tests read through the data)
    fields.push_back(arrow::field(arg, r[0]->type()));
    arrays.push_back(r[0]);
  }
  return DataFrame(arrow::schema(fields), static_cast<int64_t>(groupSize()), arrays, with_index ? unique() : nullptr);
}

arrow::Result<Series> GroupBy::seriesOf(std::string const& arg, uint32_t bit, bool drop_validity, bool with_index) {
  ARROW_ASSIGN_OR_RAISE(auto r, aggregate(arg, bit, drop_validity));
  return Series(r[0], with_index ? unique() : nullptr, arg);
}

// GROUPBY_NUMERIC_AGG(mean,double) / (count,int64_t): dataframe.cpp:1512,1526
arrow::Result<DataFrame> GroupBy::mean(std::vector<std::string> const& args) { return frameOf(args, kMean, true, true); }
arrow::Result<Series> GroupBy::mean(std::string const& arg) { return seriesOf(arg, kMean, true, true); }
arrow::Result<DataFrame> GroupBy::count(std::vector<std::string> const& args) { return frameOf(args, kCount, true, true); }
arrow::Result<Series> GroupBy::count(std::string const& arg) { return seriesOf(arg, kCount, true, true); }
// GROUPBY_AGG(max|min|sum): dataframe.cpp:1530-1534
arrow::Result<DataFrame> GroupBy::max(std::vector<std::string> const& args) { return frameOf(args, kMax, false, true); }
arrow::Result<Series> GroupBy::max(std::string const& arg) { return seriesOf(arg, kMax, false, true); }
arrow::Result<DataFrame> GroupBy::min(std::vector<std::string> const& args) { return frameOf(args, kMin, false, true); }
arrow::Result<Series> GroupBy::min(std::string const& arg) { return seriesOf(arg, kMin, false, true); }
arrow::Result<DataFrame> GroupBy::sum(std::vector<std::string> const& args) { return frameOf(args, kSum, false, true); }
arrow::Result<Series> GroupBy::sum(std::string const& arg) { return seriesOf(arg, kSum, false, true); }
// GROUPBY_AGG(product) dataframe.cpp:1536; GROUPBY_NUMERIC_AGG(stddev|variance, double) :1516,1520 (validity dropped)
arrow::Result<DataFrame> GroupBy::product(std::vector<std::string> const& args) { return frameOf(args, PA_AGG_PRODUCT, false, true); }
arrow::Result<Series> GroupBy::product(std::string const& arg) { return seriesOf(arg, PA_AGG_PRODUCT, false, true); }
arrow::Result<DataFrame> GroupBy::variance(std::vector<std::string> const& args) { return frameOf(args, PA_AGG_VARIANCE, true, true); }
arrow::Result<Series> GroupBy::variance(std::string const& arg) { return seriesOf(arg, PA_AGG_VARIANCE, true, true); }
arrow::Result<DataFrame> GroupBy::stddev(std::vector<std::string> const& args) { return frameOf(args, PA_AGG_STDDEV, true, true); }
arrow::Result<Series> GroupBy::stddev(std::string const& arg) { return seriesOf(arg, PA_AGG_STDDEV, true, true); }
// GROUPBY_NUMERIC_AGG(all|any, bool) dataframe.cpp:1522,1524
arrow::Result<DataFrame> GroupBy::all(std::vector<std::string> const& args) { return frameOf(args, PA_AGG_BOOL_ALL, true, true); }
arrow::Result<Series> GroupBy::all(std::string const& arg) { return seriesOf(arg, PA_AGG_BOOL_ALL, true, true); }
arrow::Result<DataFrame> GroupBy::any(std::vector<std::string> const& args) { return frameOf(args, PA_AGG_BOOL_ANY, true, true); }
arrow::Result<Series> GroupBy::any(std::string const& arg) { return seriesOf(arg, PA_AGG_BOOL_ANY, true, true); }
// GROUPBY_NUMERIC_AGG(count_distinct, int64_t) dataframe.cpp:1528
arrow::Result<DataFrame> GroupBy::count_distinct(std::vector<std::string> const& args) { return frameOf(args, PA_AGG_COUNT_DISTINCT, true, true); }
arrow::Result<Series> GroupBy::count_distinct(std::string const& arg) { return seriesOf(arg, PA_AGG_COUNT_DISTINCT, true, true); }
// dataframe.cpp:1698-1806: first(vector) is indexed by the keys, first(string) is not (:1748),
// last(vector) is not (:1781), last(string) is (:1805)
arrow::Result<DataFrame> GroupBy::first(std::vector<std::string> const& args) { return frameOf(args, kFirst, false, true); }
arrow::Result<Series> GroupBy::first(std::string const& arg) { return seriesOf(arg, kFirst, false, false); }
arrow::Result<DataFrame> GroupBy::last(std::vector<std::string> const& args) { return frameOf(args, kLast, false, false); }
arrow::Result<Series> GroupBy::last(std::string const& arg) { return seriesOf(arg, kLast, false, true); }

// dataframe.cpp:1602-1696: one pass, two columns; no key index
arrow::Result<DataFrame> GroupBy::min_max(std::vector<std::string> const& args) {
  arrow::FieldVector fields;
  arrow::ArrayVector arrays;
  for (auto const& arg : args) {
    ARROW_ASSIGN_OR_RAISE(auto r, aggregate(arg, kMin | kMax, false));
    fields.push_back(arrow::field(arg + "_min", r[0]->type()));
    fields.push_back(arrow::field(arg + "_max", r[1]->type()));
    arrays.push_back(r[0]);
    arrays.push_back(r[1]);
  }
  return DataFrame(arrow::schema(fields), static_cast<int64_t>(groupSize()), arrays);
}
arrow::Result<DataFrame> GroupBy::min_max(std::string const& arg) {
  ARROW_ASSIGN_OR_RAISE(auto r, aggregate(arg, kMin | kMax, false));
  return DataFrame(arrow::schema({arrow::field("min", r[0]->type()), arrow::field("max", r[1]->type())}),
                   static_cast<int64_t>(groupSize()), {r[0], r[1]});
}

// ------------------------------ Resampler ------------------------------
Resampler::Resampler(DataFrame const& _df) : GroupBy("__resampler_idx__", _df) {}

Resampler::Resampler(DataFrame const& _df, int64_t freq_ns, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                     int64_t offset_ns) {
  df = _df;
  key_array = df.indexArray();
  if (!key_array) throw std::runtime_error("frame has no index");
  Exported k(*key_array);
  pa_options opt;
  pa_options_init(&opt);
  if (pa_resample_create(&k.dev, &k.schema, freq_ns, closed_right, label_right, origin_code(origin.type), origin.custom_ns,
                         offset_ns, &opt, &handle) != PA_OK)
    throw_pa("resample");
}

Resampler::Resampler(DataFrame const& _df, DownsampleRule const& rule) {
  df = _df;
  key_array = df.indexArray();
  if (!key_array) throw std::runtime_error("frame has no index");
  Exported k(*key_array);
  pa_options opt;
  pa_options_init(&opt);
  if (pa_downsample_create(&k.dev, &k.schema, rule.multiple, rule.unit, rule.closed_label_right, rule.week_starts_monday,
                           rule.calendar_based_origin, &opt, &handle) != PA_OK)
    throw_pa("downsample");
}

arrow::Result<DataFrame> Resampler::frameOfAll(std::string const& name) {
  // RESAMPLE_GROUP_BY_FUNCTION (group_by.h:249-253): GroupBy::name(all columns)->setIndex(unique())
  const auto cols = getDF().columnNames();
  arrow::Result<DataFrame> r = arrow::Status::NotImplemented(name);
  if (name == "mean") r = GroupBy::mean(cols);
  else if (name == "count") r = GroupBy::count(cols);
  else if (name == "max") r = GroupBy::max(cols);
  else if (name == "min") r = GroupBy::min(cols);
  else if (name == "sum") r = GroupBy::sum(cols);
  else if (name == "first") r = GroupBy::first(cols);
  else if (name == "last") r = GroupBy::last(cols);
  else if (name == "product") r = GroupBy::product(cols);
  else if (name == "variance") r = GroupBy::variance(cols);
  else if (name == "stddev") r = GroupBy::stddev(cols);
  ARROW_RETURN_NOT_OK(r.status());
  return r->setIndex(unique());
}

std::optional<DateOffset> DateOffset::FromString(std::string const& code) {
  const auto [unit, mul] = splitTimeSpan(code);
  DateOffset o;
  o.multiplier = mul;
  if (unit == "D") o.type = Day;
  else if (unit == "WS") o.type = WeekStart;
  else if (unit == "W") o.type = WeekEnd;
  else if (unit == "MS") o.type = MonthStart;
  else if (unit == "M") o.type = MonthEnd;
  else if (unit == "Y") o.type = YearEnd;
  else if (unit == "YS") o.type = YearStart;
  else if (unit == "Q") o.type = QuarterEnd;
  else if (unit == "QS") o.type = QuarterStart;
  else return std::nullopt;
  return o;
}

Resampler::Resampler(DataFrame const& _df, DateOffset const& rule, bool closed_right, bool label_right) {
  df = _df;
  key_array = df.indexArray();
  if (!key_array) throw std::runtime_error("frame has no index");
  Exported k(*key_array);
  pa_options opt;
  pa_options_init(&opt);
  if (pa_resample_create_calendar(&k.dev, &k.schema, static_cast<int32_t>(rule.type), rule.multiplier, closed_right, label_right,
                                  &opt, &handle) != PA_OK)
    throw_pa("resample");
}

namespace {
// resample.h:61-86: minute / second / ... units are fixed-width rules, everything else goes to DateOffset::FromString
std::optional<time_duration> rule_to_duration(std::string const& rule) {
  auto [unit, value] = splitTimeSpan(rule);
  if (unit == "T" || unit == "min") return minutes(value);
  if (unit == "S") return seconds(value);
  if (unit == "L" || unit == "ms") return milliseconds(value);
  if (unit == "U" || unit == "us") return microseconds(value);
  if (unit == "N" || unit == "ns") return nanoseconds(value);
  return std::nullopt;
}
DateOffset rule_to_offset(std::string const& rule) {
  auto o = DateOffset::FromString(rule);
  if (!o) throw std::runtime_error("Invalid time offset " + rule);   // (the reference dereferences an empty optional here)
  return *o;
}
}  // namespace

Resampler resample(DataFrame const& df, time_duration const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                   time_duration const& offset, std::string const& tz) {
  if (!tz.empty()) throw std::runtime_error("resample: tz is not supported");
  return Resampler(df, rule.count(), closed_right, label_right, origin, offset.count());
}
Resampler resample(DataFrame const& df, DateOffset const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const&,
                   time_duration const&, std::string const& tz) {
  // (origin and offset only enter adjustDatesAnchored, which the DateOffset branch never calls: resample.cpp:248-267)
  if (!tz.empty()) throw std::runtime_error("resample: tz is not supported");
  return Resampler(df, rule, closed_right, label_right);
}
Resampler resample(DataFrame const& df, std::string const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                   time_duration const& offset, std::string const& tz) {
  if (auto d = rule_to_duration(rule)) return resample(df, *d, closed_right, label_right, origin, offset, tz);
  return resample(df, rule_to_offset(rule), closed_right, label_right, origin, offset, tz);
}
Resampler resample(Series const& s, time_duration const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                   time_duration const& offset, std::string const& tz) {
  // resample.h:109-115
  DataFrame df(arrow::schema({arrow::field(s.name(), s.dtype())}), s.size(), {s.array()}, s.indexArray());
  return resample(df, rule, closed_right, label_right, origin, offset, tz);
}
Resampler resample(Series const& s, DateOffset const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                   time_duration const& offset, std::string const& tz) {
  DataFrame df(arrow::schema({arrow::field(s.name(), s.dtype())}), s.size(), {s.array()}, s.indexArray());
  return resample(df, rule, closed_right, label_right, origin, offset, tz);
}
Resampler resample(Series const& s, std::string const& rule, bool closed_right, bool label_right, TimeGrouperOrigin const& origin,
                   time_duration const& offset, std::string const& tz) {
  if (auto d = rule_to_duration(rule)) return resample(s, *d, closed_right, label_right, origin, offset, tz);
  return resample(s, rule_to_offset(rule), closed_right, label_right, origin, offset, tz);
}

// ------------------------------ sort, ingest ------------------------------
namespace {
struct SortHandle {
  pa_groupby* h = nullptr;
  SortHandle(const arrow::Array& values, bool ascending) {
    Exported v(values);
    pa_options opt;
    pa_options_init(&opt);
    if (pa_sort_create(&v.dev, &v.schema, ascending, &opt, &h) != PA_OK) throw_pa("sort");
  }
  ~SortHandle() { pa_groupby_destroy(h); }
  ArrayPtr indices() const {
    ArrowArray a;
    ArrowSchema s;
    if (pa_sort_indices(h, &a, &s) != PA_OK) throw_pa("pa_sort_indices");
    return ReturnOrThrowOnFailure(import_result(&a, &s));
  }
  ArrayPtr take(const ArrayPtr& col, const ArrayPtr& idx) const {
    if (device_takeable(*col->type())) {
      Exported v(*col);
      ArrowArray a;
      ArrowSchema s;
      if (pa_groupby_take_grouped(h, &v.dev, &v.schema, &a, &s) != PA_OK) throw_pa("pa_groupby_take_grouped");
      auto out = ReturnOrThrowOnFailure(import_result(&a, &s));
      if (!out->type()->Equals(col->type())) out = ReturnOrThrowOnFailure(out->View(col->type()));
      return out;
    }
    return ReturnOrThrowOnFailure(arrow::compute::Take(*col, *idx));   // strings / nested: host take on the device-made order
  }
};
}  // namespace

Series Series::argsort(bool ascending) const {
  SortHandle sh(*m_array, ascending);
  return Series(sh.indices(), m_index, m_name);
}

Series Series::sort(bool ascending) const {
  if (!m_index) throw std::runtime_error("Cannot sort a Series without an index");   // series.cpp:979-982
  SortHandle sh(*m_array, ascending);
  ArrayPtr idx;
  if (!device_takeable(*m_array->type()) || !device_takeable(*m_index->type())) idx = sh.indices();
  return Series(sh.take(m_array, idx), sh.take(m_index, idx), m_name);
}

DataFrame DataFrame::sort_index(bool ascending, bool ignore_index) const {
  SortHandle sh(*m_index, ascending);
  ArrayPtr idx;
  for (auto const& c : m_array->columns()) if (!device_takeable(*c->type())) { idx = sh.indices(); break; }
  arrow::ArrayVector cols;
  for (auto const& c : m_array->columns()) cols.push_back(sh.take(c, idx));
  auto rb = arrow::RecordBatch::Make(m_array->schema(), m_array->num_rows(), cols);
  return DataFrame(rb, ignore_index ? nullptr : sh.take(m_index, idx));
}

DataFrame DataFrame::sort_values(std::vector<std::string> const& by, bool ascending) const {
  auto array = m_array;
  for (auto const& field : by) {
    const int i = m_array->schema()->GetFieldIndex(field);
    if (i < 0) throw std::runtime_error(field + " not in schema");
    auto col = m_array->column(i);
    array = ReturnOrThrowOnFailure(array->SetColumn(i, arrow::field(field, col->type()), Series(col, m_index).sort(ascending).array()));
  }
  return DataFrame(array);
}

DataFrame DataFrame::readBinary(std::basic_string_view<uint8_t> const& blob, std::optional<std::string> const& index) {
  auto buffer = std::make_shared<arrow::Buffer>(blob.data(), static_cast<int64_t>(blob.size()));
  auto reader = ReturnOrThrowOnFailure(arrow::ipc::RecordBatchStreamReader::Open(std::make_shared<arrow::io::BufferReader>(buffer)));
  auto batches = ReturnOrThrowOnFailure(reader->ToRecordBatches());
  if (batches.size() != 1)
    throw std::invalid_argument("PandasArrow Cannot ReadBinary from a Table or Array of RecordBatches yet. Always Assume Single RecordBatch.");
  ArrayPtr indexPtr;
  if (index) {
    const int pos = batches[0]->schema()->GetFieldIndex(*index);
    if (pos != -1) {
      indexPtr = batches[0]->column(pos);
      if (indexPtr->type_id() == arrow::Type::INT64) indexPtr = ReturnOrThrowOnFailure(indexPtr->View(arrow::timestamp(arrow::TimeUnit::NANO)));
      batches[0] = ReturnOrThrowOnFailure(batches[0]->RemoveColumn(pos));
    }
  }
  return DataFrame(batches[0], indexPtr);
}

DataFrame DataFrame::readParquet(std::string const& path) {
  auto infile = ReturnOrThrowOnFailure(arrow::io::ReadableFile::Open(path));
  auto reader = ReturnOrThrowOnFailure(parquet::arrow::OpenFile(infile, arrow::default_memory_pool()));
  std::shared_ptr<arrow::Table> table;
  ThrowOnFailure(reader->ReadTable(&table));
  arrow::TableBatchReader tbr(*table);
  auto rb = ReturnOrThrowOnFailure(tbr.ToRecordBatches());
  if (rb.empty()) throw std::runtime_error("Cannot Initialize DataFrame with empty parquet table");
  if (rb.size() != 1)
    throw std::runtime_error("DataFrame Only supports Parquet Table with single record batch\nFound " + std::to_string(rb.size()) + " record batches\n");
  return DataFrame(rb[0]);
}

DataFrame DataFrame::to_device() const {
  auto set = std::make_shared<DeviceColumns>();
  auto ingest = [&](const ArrayPtr& a) {
    if (!a || a->length() == 0) return;
    const auto id = a->type_id();
    const bool str = id == arrow::Type::STRING || id == arrow::Type::LARGE_STRING;
    if (!str && !device_takeable(*a->type())) return;                        // nested / dictionary columns stay on the host
    if (find_device_copy(*a)) return;
    ArrowDeviceArray host{};
    ArrowSchema sc{};
    ThrowOnFailure(arrow::ExportArray(*a, &host.array, &sc));
    host.device_id = -1;
    host.device_type = ARROW_DEVICE_CPU;
    auto col = std::make_shared<DeviceColumn>();
    pa_options opt;
    pa_options_init(&opt);
    const int rc = pa_column_to_device(&host, &sc, &opt, &col->dev);
    host.array.release(&host.array);
    sc.release(&sc);
    if (rc != PA_OK) throw_pa("to_device");
    col->host = a->data();
    {
      std::lock_guard<std::mutex> lock(g_device_mutex);
      g_device_registry[device_key(*a->data())] = col;
    }
    set->cols.push_back(std::move(col));
  };
  ingest(m_index);
  if (m_array) for (auto const& c : m_array->columns()) ingest(c);
  DataFrame out(*this);
  out.m_device = set;
  return out;
}

// ------------------------------ helpers ------------------------------
ArrayPtr range(int64_t start, int64_t end) {
  arrow::Int64Builder b;
  ThrowOnFailure(b.Reserve(std::max<int64_t>(end - start, 0)));
  for (int64_t i = start; i < end; ++i) b.UnsafeAppend(i);
  return ReturnOrThrowOnFailure(b.Finish());
}

ArrayPtr date_range(int64_t start_ns, int periods, time_duration freq) {
  arrow::TimestampBuilder b(arrow::timestamp(arrow::TimeUnit::NANO), arrow::default_memory_pool());
  for (int i = 0; i < periods; ++i) ThrowOnFailure(b.Append(start_ns + i * freq.count()));
  return ReturnOrThrowOnFailure(b.Finish());
}

int64_t ns_from_ymd(int y, int m, int d) {
  // days from civil (Howard Hinnant)
  y -= m <= 2;
  const int64_t era = (y >= 0 ? y : y - 399) / 400;
  const unsigned yoe = static_cast<unsigned>(y - era * 400);
  const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return (era * 146097 + static_cast<int64_t>(doe) - 719468) * 86400LL * 1000000000LL;
}

}  // namespace pd
