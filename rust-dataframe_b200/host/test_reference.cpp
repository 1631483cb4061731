// test_reference.cpp -- the reference's own unit tests for the hot path, re-stated against the C++ host
// mirror (functions.hpp) so they read like the originals; every value comes from the CUDA kernels through
// the C ABI.  Test names and literals follow
//   reference src/functions/scalar.rs:558-671, src/functions/aggregate.rs:105-147,
//   src/operation/scalar.rs:320-342 (the Cast-then-Add plan), src/dataframe.rs:783-808.
// Needs a B200; run by tests/test_host_mirror.py (-m gpu).
#include <cmath>
#include <cstdio>
#include <limits>

#include "functions.hpp"

using namespace rdf;

static int g_failed = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); g_failed++; } } while (0)
#define RUN(fn) do { std::printf("test %s ...\n", #fn); int before = g_failed; fn(); std::printf("test %s ... %s\n", #fn, g_failed == before ? "ok" : "FAILED"); } while (0)

static const double EPS = std::numeric_limits<double>::epsilon();

static void test_primitive_array_abs_f64() {  // scalar.rs:565-573
    auto a = Float64Array::from(std::vector<double>{-5.2, -6.1, 7.3, -8.6, -0.0});
    auto r = ScalarFunctions::abs<double>({&a});
    const auto& c = r.unwrap()[0];
    CHECK(5.2 - c.value(0) < EPS); CHECK(6.1 - c.value(1) < EPS); CHECK(7.3 - c.value(2) < EPS);
    CHECK(8.6 - c.value(3) < EPS); CHECK(0.0 - c.value(4) < EPS);
    CHECK(c.value(0) == 5.2 && c.value(3) == 8.6 && !std::signbit(c.value(4)));  // two-sided
}

static void test_primitive_array_abs_i32() {  // scalar.rs:576-584
    auto a = Int32Array::from(std::vector<int32_t>{-5, -6, 7, -8, -0});
    auto r = ScalarFunctions::abs<int32_t>({&a});
    const auto& c = r.unwrap()[0];
    CHECK(5 == c.value(0)); CHECK(6 == c.value(1)); CHECK(7 == c.value(2)); CHECK(8 == c.value(3)); CHECK(0 == c.value(4));
}

static void test_primitive_array_acos_f64() {  // scalar.rs:587-593
    auto a = Float64Array::from(std::vector<double>{-0.2, 0.25, 0.75});
    auto r = ScalarFunctions::acos<double>({&a});
    const auto& c = r.unwrap()[0];
    const double want[3] = {1.7721542475852274, 1.318116071652818, 0.7227342478134157};
    for (int i = 0; i < 3; i++) { CHECK(want[i] - c.value(i) < EPS); CHECK(std::fabs(want[i] - c.value(i)) <= 4 * EPS * want[i]); }
}

static void test_primitive_array_cos_f64() {  // scalar.rs:596-602
    auto a = Float64Array::from(std::vector<double>{-0.2, 0.25, 0.75});
    auto r = ScalarFunctions::cos<double>({&a});
    const auto& c = r.unwrap()[0];
    const double want[3] = {0.9800665778412416, 0.9689124217106447, 0.7316888688738209};
    for (int i = 0; i < 3; i++) { CHECK(want[i] - c.value(i) < EPS); CHECK(std::fabs(want[i] - c.value(i)) <= 3 * EPS * want[i]); }
}

static void test_aggregate_count() {  // aggregate.rs:123-127
    auto a = Int32Array::from(std::vector<int32_t>{5, 6, 7, 8, 9});
    auto c = AggregateFunctions::count<int32_t>({&a}).value();
    CHECK(5 == c);
}

static void test_aggregate_mean() {  // aggregate.rs:130-146
    auto a = Int32Array::from(std::vector<int32_t>{0, 1, 2, 3, 4});
    auto b = Int32Array::from(std::vector<int32_t>{5, 6, 7, 8, 9});
    auto c = AggregateFunctions::avg<int32_t>({&a, &b});
    CHECK(c.has_value() && *c == 4.5);
    auto d = Int32Array::from(std::vector<std::optional<int32_t>>{0, std::nullopt, 1, std::nullopt, 2, 3, 4});
    auto e = AggregateFunctions::avg<int32_t>({&d, &b});
    CHECK(e.has_value() && *e == 4.5);
}

static void bench_multiply_i32_input() {  // scalar.rs:621-671: 380 chunks of [None, 200, None, -256, None]
    std::vector<Int32Array> owned;
    for (int i = 0; i < 380; i++) owned.push_back(Int32Array::from(std::vector<std::optional<int32_t>>{std::nullopt, 200, std::nullopt, -256, std::nullopt}));
    std::vector<const Int32Array*> chunks;
    for (auto& a : owned) chunks.push_back(&a);
    auto r = ScalarFunctions::par_multiply<int32_t>(chunks, chunks);
    auto& out = r.unwrap();
    CHECK(out.size() == 380);
    for (auto& c : out) {
        CHECK(c.len() == 5 && c.null_count() == 3);
        CHECK(c.is_null(0) && c.is_null(2) && c.is_null(4));
        CHECK(c.value(1) == 40000 && c.value(3) == 65536);
    }
}

static void test_dataframe_ops_row0() {  // dataframe.rs:783-808: lat + lng on the first rows of the CSV fixture
    auto lat = Float64Array::from(std::vector<double>{57.653484, 53.002666, 52.412811, 51.481583});
    auto lng = Float64Array::from(std::vector<double>{-3.335724, -2.179404, -1.778197, -3.179090});
    auto r = ScalarFunctions::add<double>({&lat}, {&lng});
    const auto& c = r.unwrap()[0];
    CHECK(54.31776 - c.value(0) < 0.0001);
    CHECK(c.value(0) == 57.653484 + -3.335724);  // IEEE add is exact: bit-identical to the host
    CHECK(c.null_count() == 0 && c.len() == 4);
    auto s = AggregateFunctions::sum<double>({&c});
    CHECK(s.has_value() && std::fabs(*s - (c.value(0) + c.value(1) + c.value(2) + c.value(3))) < 1e-12);
}

static void test_add_operation_plan_cast_then_add() {  // operation/scalar.rs:337-340: [Cast b -> Int64, Add]
    auto a = Int64Array::from(std::vector<int64_t>{1, 2, 3, 4000000000LL});
    auto b = Int32Array::from(std::vector<std::optional<int32_t>>{10, std::nullopt, -30, 40});
    auto b64 = cast<int32_t, int64_t>({&b});
    auto& bc = b64.unwrap();
    auto r = ScalarFunctions::add<int64_t>({&a}, {&bc[0]});
    const auto& c = r.unwrap()[0];
    CHECK(c.value(0) == 11 && c.is_null(1) && c.value(2) == -27 && c.value(3) == 4000000040LL && c.null_count() == 1);
}

static void test_errors() {
    auto a = Int32Array::from(std::vector<int32_t>{6, 8, 10});
    auto z = Int32Array::from(std::vector<int32_t>{2, 0, 5});
    auto r = ScalarFunctions::divide<int32_t>({&a}, {&z});
    CHECK(r.is_err() && r.unwrap_err().kind == ArrowError::DivideByZero);  // arrow: Err(ArrowError::DivideByZero)
    auto zn = Int32Array::from(std::vector<std::optional<int32_t>>{2, std::nullopt, 5});
    auto ok = ScalarFunctions::divide<int32_t>({&a}, {&zn});
    CHECK(ok.is_ok() && ok.unwrap()[0].value(0) == 3 && ok.unwrap()[0].is_null(1) && ok.unwrap()[0].value(2) == 2);
    auto fz = Float64Array::from(std::vector<double>{1.0, 0.0});
    auto fa = Float64Array::from(std::vector<double>{1.0, 1.0});
    CHECK(ScalarFunctions::divide<double>({&fa}, {&fz}).is_err());          // floats too
    auto s = Int32Array::from(std::vector<int32_t>{1, 2});
    auto m = ScalarFunctions::add<int32_t>({&a}, {&s});
    CHECK(m.is_err() && m.unwrap_err().message == "Cannot perform math operation on arrays of different length");  // scalar.rs:508-511
    bool panicked = false;
    auto alln = Int32Array::from(std::vector<std::optional<int32_t>>{std::nullopt, std::nullopt});
    try { AggregateFunctions::max<int32_t>({&a, &alln}); } catch (const ReferencePanic&) { panicked = true; }
    CHECK(panicked);  // compute::max(..).unwrap() on None (aggregate.rs:19)
    CHECK(!AggregateFunctions::max<int32_t>({}).has_value());
    CHECK(AggregateFunctions::sum<int32_t>({&a, &alln}).value() == 24);
    CHECK(AggregateFunctions::max<int32_t>({&a}).value() == 10 && AggregateFunctions::min<int32_t>({&a}).value() == 6);
    auto sl = a.slice(1, 2);
    CHECK(AggregateFunctions::sum<int32_t>({&sl}).value() == 18);
}

int main() {
    RUN(test_primitive_array_abs_f64);
    RUN(test_primitive_array_abs_i32);
    RUN(test_primitive_array_acos_f64);
    RUN(test_primitive_array_cos_f64);
    RUN(test_aggregate_count);
    RUN(test_aggregate_mean);
    RUN(bench_multiply_i32_input);
    RUN(test_dataframe_ops_row0);
    RUN(test_add_operation_plan_cast_then_add);
    RUN(test_errors);
    std::printf("%s (%d failed checks)\n", g_failed ? "FAILED" : "ALL OK", g_failed);
    return g_failed ? 1 : 0;
}
