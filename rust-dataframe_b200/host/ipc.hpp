// ipc.hpp -- C++ face of the Arrow IPC entries of the C ABI (DataFrame::from_arrow / to_arrow, src/dataframe.rs:391-407,
// 515-525).  RAII around bdf_ipc; the file is mapped, views point into the mapping and live as long as the IpcFile.
#pragma once

#include <string>
#include <utility>
#include <vector>

#include "b200df.h"
#include "primitive_array.hpp"

namespace rdf {

struct IpcField {
    std::string name;
    int dtype;        // bdf_dtype, BDF_BOOL, or -1 for a column type outside the numeric path
    bool nullable;
};

class IpcFile {
   public:
    explicit IpcFile(const std::string& path) {
        if (bdf_ipc_open(path.c_str(), &f_) != BDF_OK) throw std::runtime_error(std::string("from_arrow: ") + bdf_last_error());
        int32_t nc = 0;
        bdf_ipc_describe(f_, &nc, &n_batches_, &n_rows_);
        for (int32_t c = 0; c < nc; c++) {
            const char* name = nullptr;
            int32_t dt = -1, nl = 0;
            bdf_ipc_column(f_, c, &name, &dt, &nl);
            fields_.push_back({name, dt, nl != 0});
        }
    }
    ~IpcFile() { bdf_ipc_close(f_); }
    IpcFile(const IpcFile&) = delete;
    IpcFile& operator=(const IpcFile&) = delete;

    const std::vector<IpcField>& schema() const { return fields_; }
    int64_t num_batches() const { return n_batches_; }
    int64_t num_rows() const { return n_rows_; }
    int64_t batch_rows(int64_t b) const {
        int64_t r = 0;
        if (bdf_ipc_batch_rows(f_, b, &r) != BDF_OK) throw std::out_of_range(bdf_last_error());
        return r;
    }
    // Zero-copy view of one column of one RecordBatch (host pointers into the mapping).
    bdf_view view(int64_t batch, int32_t col) const {
        bdf_view v{};
        if (bdf_ipc_view(f_, batch, col, &v) != BDF_OK) throw std::runtime_error(bdf_last_error());
        return v;
    }
    // The chosen columns on the device, one chunk per RecordBatch.
    std::vector<bdf_col*> read(bdf_ctx* ctx, const std::vector<int32_t>& cols, int flags = 0) const {
        std::vector<bdf_col*> out(cols.size(), nullptr);
        if (bdf_ipc_read(ctx, f_, (int32_t)cols.size(), cols.data(), flags, out.data()) != BDF_OK) throw std::runtime_error(bdf_last_error());
        return out;
    }
    const bdf_ipc* handle() const { return f_; }

   private:
    bdf_ipc* f_ = nullptr;
    std::vector<IpcField> fields_;
    int64_t n_batches_ = 0, n_rows_ = 0;
};

// Host arrays -> IPC file: columns[c][b] is column c of RecordBatch b.
template <typename... Ts>
inline void write_ipc_host(const std::string& path, const std::vector<std::string>& names, const std::vector<int32_t>& dtypes,
                           const std::vector<std::vector<bdf_view>>& columns) {
    std::vector<const char*> cn;
    std::vector<const bdf_view*> cp;
    for (auto& s : names) cn.push_back(s.c_str());
    for (auto& c : columns) cp.push_back(c.data());
    const int64_t nb = columns.empty() ? 0 : (int64_t)columns[0].size();
    if (bdf_ipc_write_host(path.c_str(), (int32_t)names.size(), cn.data(), dtypes.data(), nb, cp.data()) != BDF_OK)
        throw std::runtime_error(std::string("to_arrow: ") + bdf_last_error());
}

}  // namespace rdf
