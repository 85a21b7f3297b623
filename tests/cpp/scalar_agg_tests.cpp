// Whole-column aggregates of the façade (NDFrame::mean/min/max/count/min_max/sum,
// /root/reference/src/ndframe.cpp:119,160-175,220) — computed on the GPU as ONE group through the C ABI — against
// the arrow::compute scalar kernels the reference calls for them.  Separate from facade_tests so that it runs (and can
// fail) on its own.  Needs a GPU (run by tests/test_zz_facade_scalar_gpu.py).
#include <cmath>
#include <cstdio>
#include <random>

#include <arrow/compute/api.h>

#include "../../pandasarrow_b200/csrc/host/pd_groupby.h"

static int g_fail = 0, g_checks = 0;
#define REQUIRE(cond)                                                              \
  do {                                                                             \
    ++g_checks;                                                                    \
    if (!(cond)) { ++g_fail; std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); } \
  } while (0)

namespace ac = arrow::compute;

static std::shared_ptr<arrow::Scalar> host(const char* fn, const pd::ArrayPtr& a, bool skip_null = true) {
  ac::ScalarAggregateOptions opt{skip_null};
  return pd::ReturnOrThrowOnFailure(ac::CallFunction(fn, {a}, &opt)).scalar();
}

static bool same(const pd::Scalar& got, const std::shared_ptr<arrow::Scalar>& want, double rtol = 0.0) {
  if (!got.scalar->type->Equals(*want->type)) { std::printf("  dtype %s vs %s\n", got.scalar->type->ToString().c_str(), want->type->ToString().c_str()); return false; }
  if (got.scalar->is_valid != want->is_valid) return false;
  if (!want->is_valid) return true;
  if (rtol > 0.0) {
    const double g = std::static_pointer_cast<arrow::DoubleScalar>(got.scalar)->value, w = std::static_pointer_cast<arrow::DoubleScalar>(want)->value;
    return std::fabs(g - w) <= rtol * std::fmax(std::fabs(w), 1e-300);
  }
  return got.scalar->Equals(*want);
}

template <class Builder, class T>
static pd::ArrayPtr make(const std::vector<T>& v, const std::vector<bool>& valid = {}) {
  Builder b;
  for (size_t i = 0; i < v.size(); ++i) {
    if (!valid.empty() && !valid[i]) pd::ThrowOnFailure(b.AppendNull());
    else pd::ThrowOnFailure(b.Append(v[i]));
  }
  return pd::ReturnOrThrowOnFailure(b.Finish());
}

static void check_column(const pd::ArrayPtr& a, bool floating) {
  pd::Series s(a, nullptr, "v");
  for (bool skip : {true, false}) {
    REQUIRE(same(s.min(skip), host("min", a, skip)));
    REQUIRE(same(s.max(skip), host("max", a, skip)));
    auto mm = s.min_max(skip);
    REQUIRE(same(mm.first, host("min", a, skip)) && same(mm.second, host("max", a, skip)));
    REQUIRE(same(s.mean(skip), host("mean", a, skip), 1e-12));
    REQUIRE(same(s.sum(skip), host("sum", a, skip), floating ? 1e-12 : 0.0));
    REQUIRE(same(s.agg("sum", skip), host("sum", a, skip), floating ? 1e-12 : 0.0));
    // NDFrame::first / last (ndframe.cpp:129,160): arrow's scalar kernels, which skip nulls unless told not to
    REQUIRE(same(s.first(skip), host("first", a, skip)));
    REQUIRE(same(s.last(skip), host("last", a, skip)));
    REQUIRE(same(s.agg("max", skip), host("max", a, skip)));
  }
  {
    // DataFrame::sum (ndframe.cpp:220 over the concatenated columns): two copies of the column
    auto rb = arrow::RecordBatch::Make(arrow::schema({arrow::field("a", a->type()), arrow::field("b", a->type())}), a->length(), {a, a});
    pd::DataFrame df(rb);
    auto chunked = std::make_shared<arrow::ChunkedArray>(arrow::ArrayVector{a, a});
    auto want = pd::ReturnOrThrowOnFailure(ac::CallFunction("sum", {chunked})).scalar();
    REQUIRE(same(df.sum(), want, floating ? 1e-12 : 0.0));
  }
  REQUIRE(s.count() == std::static_pointer_cast<arrow::Int64Scalar>(pd::ReturnOrThrowOnFailure(ac::CallFunction("count", {a})).scalar())->value);
}

int main() {
  try {
    pd::ThrowOnFailure(ac::Initialize());
    std::mt19937_64 rng(5);
    std::vector<int32_t> iv(10007);
    std::vector<double> dv(10007);
    std::vector<bool> valid(10007);
    for (size_t i = 0; i < iv.size(); ++i) { iv[i] = int32_t(rng() % 2001) - 1000; dv[i] = double(rng() % 100000) / 7.0 - 5000.0; valid[i] = rng() % 10 != 0; }
    check_column(make<arrow::Int32Builder>(iv), false);
    check_column(make<arrow::Int32Builder>(iv, valid), false);
    check_column(make<arrow::DoubleBuilder>(dv), true);
    check_column(make<arrow::DoubleBuilder>(dv, valid), true);
    check_column(make<arrow::DoubleBuilder>(dv, valid)->Slice(13, 5000), true);                 // offset into values and validity
    check_column(make<arrow::Int64Builder>(std::vector<int64_t>{}), false);                      // empty: null / 0
    check_column(make<arrow::DoubleBuilder>(std::vector<double>{1.0, 2.0}, {false, false}), true); // all null
    check_column(make<arrow::DoubleBuilder>(std::vector<double>{3.5}), true);
    {   // nulls at both ends: first / last must skip them (skip_null) or return them (positional)
      std::vector<bool> ends(10007, true);
      for (int i = 0; i < 40; ++i) { ends[i] = false; ends[10006 - i] = false; }
      check_column(make<arrow::DoubleBuilder>(dv, ends), true);
      check_column(make<arrow::Int32Builder>(iv, ends), false);
    }
  } catch (std::exception const& e) {
    std::printf("EXCEPTION: %s\n", e.what());
    return 2;
  }
  std::printf("%d checks, %d failed\n", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
