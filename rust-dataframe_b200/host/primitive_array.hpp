// primitive_array.hpp -- minimal Arrow PrimitiveArray<T> for the C++ host mirror of the reference API.
// Same memory contract as arrow-rs ArrayData (and as bdf_view in include/b200df.h): a 64-byte aligned
// values buffer, an optional LSB-first validity bitmap (1 = valid), a logical offset applied to both, len,
// cached null_count.  Buffers are shared (std::shared_ptr), so slice() is zero-copy like ArrayData::slice.
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "b200df.h"

namespace rdf {

template <typename T> struct ArrowType;
#define RDF_TYPE(T, ID) template <> struct ArrowType<T> { static constexpr int id = ID; };
RDF_TYPE(int8_t, BDF_I8) RDF_TYPE(int16_t, BDF_I16) RDF_TYPE(int32_t, BDF_I32) RDF_TYPE(int64_t, BDF_I64)
RDF_TYPE(uint8_t, BDF_U8) RDF_TYPE(uint16_t, BDF_U16) RDF_TYPE(uint32_t, BDF_U32) RDF_TYPE(uint64_t, BDF_U64)
RDF_TYPE(float, BDF_F32) RDF_TYPE(double, BDF_F64)
#undef RDF_TYPE

// arrow::buffer::MutableBuffer: 64-byte aligned, capacity rounded up to 64.
class Buffer {
   public:
    explicit Buffer(size_t bytes) : size_(bytes) {
        const size_t cap = (bytes + 63) / 64 * 64 + 64;
        data_ = static_cast<uint8_t*>(std::aligned_alloc(64, cap));
        if (!data_) throw std::bad_alloc();
        std::memset(data_, 0, cap);
    }
    ~Buffer() { std::free(data_); }
    Buffer(const Buffer&) = delete;
    Buffer& operator=(const Buffer&) = delete;
    uint8_t* data() { return data_; }
    const uint8_t* data() const { return data_; }
    size_t size() const { return size_; }

   private:
    uint8_t* data_;
    size_t size_;
};

template <typename T>
class PrimitiveArray {
   public:
    using Native = T;

    PrimitiveArray() = default;

    // Float64Array::from(vec![...])
    static PrimitiveArray from(const std::vector<T>& values) {
        PrimitiveArray a;
        a.len_ = static_cast<int64_t>(values.size());
        a.values_ = std::make_shared<Buffer>(values.size() * sizeof(T));
        if (!values.empty()) std::memcpy(a.values_->data(), values.data(), values.size() * sizeof(T));
        return a;
    }
    // Int32Array::from(vec![Some(0), None, ...])
    static PrimitiveArray from(const std::vector<std::optional<T>>& values) {
        PrimitiveArray a;
        a.len_ = static_cast<int64_t>(values.size());
        a.values_ = std::make_shared<Buffer>(values.size() * sizeof(T));
        a.validity_ = std::make_shared<Buffer>((values.size() + 7) / 8);
        T* v = reinterpret_cast<T*>(a.values_->data());
        for (size_t i = 0; i < values.size(); i++) {
            if (values[i]) { v[i] = *values[i]; a.validity_->data()[i >> 3] |= uint8_t(1u << (i & 7)); }
            else a.null_count_++;
        }
        return a;
    }
    // ArrayData::new(dtype, len, Some(null_count), null_bit_buffer, 0, [values])
    static PrimitiveArray from_buffers(std::shared_ptr<Buffer> values, std::shared_ptr<Buffer> validity, int64_t len,
                                       int64_t null_count) {
        PrimitiveArray a;
        a.values_ = std::move(values); a.validity_ = std::move(validity); a.len_ = len; a.null_count_ = null_count;
        return a;
    }

    int64_t len() const { return len_; }
    int64_t offset() const { return offset_; }
    int64_t null_count() const {
        if (null_count_ < 0) {
            int64_t n = 0;
            for (int64_t i = 0; i < len_; i++) n += is_null(i);
            null_count_ = n;
        }
        return null_count_;
    }
    bool is_valid(int64_t i) const {
        if (!validity_) return true;
        const int64_t j = offset_ + i;
        return (validity_->data()[j >> 3] >> (j & 7)) & 1;
    }
    bool is_null(int64_t i) const { return !is_valid(i); }
    T value(int64_t i) const { return reinterpret_cast<const T*>(values_->data())[offset_ + i]; }
    const T* raw_values() const { return reinterpret_cast<const T*>(values_->data()) + offset_; }

    PrimitiveArray slice(int64_t offset, int64_t length) const {
        if (offset < 0 || length < 0 || offset + length > len_) throw std::out_of_range("slice");
        PrimitiveArray a = *this;
        a.offset_ = offset_ + offset; a.len_ = length; a.null_count_ = validity_ ? -1 : 0;
        return a;
    }

    bdf_view view() const {
        bdf_view v;
        v.values = values_ ? values_->data() : nullptr;
        v.validity = validity_ ? validity_->data() : nullptr;
        v.len = len_; v.offset = offset_; v.null_count = validity_ ? null_count_ : 0;
        return v;
    }

   private:
    std::shared_ptr<Buffer> values_, validity_;
    int64_t len_ = 0, offset_ = 0;
    mutable int64_t null_count_ = 0;
};

using Int8Array = PrimitiveArray<int8_t>;
using Int16Array = PrimitiveArray<int16_t>;
using Int32Array = PrimitiveArray<int32_t>;
using Int64Array = PrimitiveArray<int64_t>;
using UInt8Array = PrimitiveArray<uint8_t>;
using UInt16Array = PrimitiveArray<uint16_t>;
using UInt32Array = PrimitiveArray<uint32_t>;
using UInt64Array = PrimitiveArray<uint64_t>;
using Float32Array = PrimitiveArray<float>;
using Float64Array = PrimitiveArray<double>;

}  // namespace rdf
