// test_ipc_host.cpp -- the IPC entries of the C ABI from C++, no GPU needed: write a two-batch frame from host
// arrays (one sliced, one with nulls), read it back through the mapping, compare.  tests/test_ipc.py checks the same
// files against pyarrow; this is the C++ consumer a Rust shim would mirror.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "ipc.hpp"

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    using namespace rdf;
    const std::string path = argc > 1 ? argv[1] : "/tmp/bdf_test_ipc_host.arrow";
    // batch 0: 5 rows, batch 1: 3 rows (a slice with offset 2 of a longer array, nulls in it)
    auto a0 = PrimitiveArray<double>::from(std::vector<double>{1.5, -2.0, 3.25, 0.0, 8.0});
    auto k0 = PrimitiveArray<int32_t>::from(std::vector<std::optional<int32_t>>{7, std::nullopt, -1, std::nullopt, 4});
    auto a1_full = PrimitiveArray<double>::from(std::vector<std::optional<double>>{9.0, 9.0, 10.0, std::nullopt, 12.0, 9.0});
    auto a1 = a1_full.slice(2, 3);
    auto k1 = PrimitiveArray<int32_t>::from(std::vector<int32_t>{100, 200, 300});
    write_ipc_host(path, {"a", "k"}, {BDF_F64, BDF_I32}, {{a0.view(), a1.view()}, {k0.view(), k1.view()}});

    IpcFile f(path);
    CHECK(f.schema().size() == 2 && f.schema()[0].name == "a" && f.schema()[0].dtype == BDF_F64 && f.schema()[1].name == "k" && f.schema()[1].dtype == BDF_I32);
    CHECK(f.num_batches() == 2 && f.num_rows() == 8 && f.batch_rows(0) == 5 && f.batch_rows(1) == 3);
    const bdf_view va0 = f.view(0, 0), vk0 = f.view(0, 1), va1 = f.view(1, 0), vk1 = f.view(1, 1);
    CHECK(va0.len == 5 && va0.validity == nullptr && va0.null_count == 0 && ((const double*)va0.values)[2] == 3.25);
    CHECK(vk0.len == 5 && vk0.validity != nullptr && vk0.null_count == 2 && (vk0.validity[0] & 0x1f) == 0x15);
    CHECK(((const int32_t*)vk0.values)[0] == 7 && ((const int32_t*)vk0.values)[4] == 4);
    CHECK(va1.len == 3 && va1.offset == 0 && va1.null_count == 1 && (va1.validity[0] & 0x7) == 0x5);   // slice re-based to bit 0
    CHECK(((const double*)va1.values)[0] == 10.0 && ((const double*)va1.values)[2] == 12.0);
    CHECK(vk1.len == 3 && vk1.validity == nullptr && ((const int32_t*)vk1.values)[1] == 200);
    CHECK(((uintptr_t)va0.values & 63) == 0 && ((uintptr_t)vk1.values & 63) == 0);   // bodies start on 64-byte file offsets
    bool threw = false;
    try { f.view(2, 0); } catch (const std::exception&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { IpcFile g("/nonexistent/dir/x.arrow"); } catch (const std::exception&) { threw = true; }
    CHECK(threw);
    std::remove(path.c_str());
    std::printf("IPC HOST OK\n");
    return 0;
}
