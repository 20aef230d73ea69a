// nornic_cuda.hpp — C++ host-side mirror of the reference's Go host layer over the C ABI (include/nornic_knn.h).
//
// The reference's host side is Go (pkg/gpu/cuda/cuda_bridge.go:379-723 and pkg/gpu/gpu.go:1224-2454); no Go toolchain
// exists in the build image, so the compiled-language host mirror is C++ (header-only, links libnornic_knn.so).
// Same type / method names, argument meaning and error behaviour as the Go code, so tests/cpp/host_mirror_test.cpp
// reads like pkg/gpu/cuda/cuda_test.go and pkg/gpu/gpu_test.go.  Go's (value, error) returns become exceptions
// carrying the same sentinel identities (ErrCUDANotAvailable, ErrDeviceCreation, ErrBufferCreation,
// ErrKernelExecution, ErrInvalidBuffer — cuda_bridge.go:387-393; ErrInvalidDimensions — gpu.go).
//
//   namespace nornic::cuda   Device / Buffer / SearchResult / IsAvailable / DeviceCount / NewDevice
//   namespace nornic::gpu    EmbeddingIndex (Add / AddBatch / Remove / Search / SearchBatch / ScoreSubset /
//                            Count / Has / Get / Serialize / Deserialize), device corpus kept in step incrementally
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/nornic_knn.h"

namespace nornic {
namespace cuda {

struct CudaError : std::runtime_error { using std::runtime_error::runtime_error; };
struct ErrCUDANotAvailable : CudaError { ErrCUDANotAvailable() : CudaError("cuda: CUDA is not available on this system") {} };
struct ErrDeviceCreation : CudaError { using CudaError::CudaError; };
struct ErrBufferCreation : CudaError { using CudaError::CudaError; };
struct ErrKernelExecution : CudaError { using CudaError::CudaError; };
struct ErrInvalidBuffer : CudaError { using CudaError::CudaError; };

enum MemoryType { MemoryDevice = 0, MemoryPinned = 1 };  // cuda_bridge.go:396-404
enum class Metric { Cosine = NK_METRIC_COSINE, Dot = NK_METRIC_DOT, Euclidean = NK_METRIC_EUCLIDEAN };  // vectorspace/registry.go:27-31

struct SearchResult {  // cuda_bridge.go:425-428
    uint32_t Index;
    float Score;
};

inline std::string takeError() {  // cuda_bridge.go:452-453: read then clear
    const char *m = cuda_get_last_error();
    std::string s = m ? m : "";
    cuda_clear_error();
    return s;
}

inline bool IsAvailable() { return cuda_is_available() != 0; }                       // cuda_bridge.go:431
inline int DeviceCount() { int c = cuda_get_device_count(); return c < 0 ? 0 : c; }  // cuda_bridge.go:436

class Device;

class Buffer {  // cuda_bridge.go:417-422
public:
    Buffer(CudaBuffer *p, uint64_t size, Device *dev) : ptr_(p), size_(size), device_(dev) {}
    Buffer(const Buffer &) = delete;
    Buffer &operator=(const Buffer &) = delete;
    ~Buffer() { Release(); }
    void Release() {  // cuda_bridge.go:560-565 (idempotent)
        if (index_) { nk_index_release(index_); index_ = nullptr; }
        if (ptr_) { cuda_release_buffer(ptr_); ptr_ = nullptr; }
    }
    uint64_t Size() const { return size_; }
    std::vector<float> ReadFloat32(int count) const {  // cuda_bridge.go:573-585 (empty = Go's nil)
        if (count <= 0 || (uint64_t)count * 4 > size_ || !ptr_) return {};
        std::vector<float> out((size_t)count);
        if (cuda_buffer_copy_to_host(ptr_, out.data(), (size_t)count) != 0) return {};
        return out;
    }
    CudaBuffer *raw() const { return ptr_; }

private:
    friend class Device;
    CudaBuffer *ptr_;
    uint64_t size_;
    Device *device_;
    NkIndex *index_ = nullptr;  // cached fused-search handle over this buffer
    uint64_t index_n_ = 0;
    uint32_t index_dim_ = 0;
    int index_metric_ = -1;
};

class Device {  // cuda_bridge.go:407-415
public:
    explicit Device(int deviceID) : id_(deviceID) {
        if (!IsAvailable()) throw ErrCUDANotAvailable();
        ptr_ = cuda_create_device(deviceID);
        if (!ptr_) throw ErrDeviceCreation("cuda: failed to create CUDA device: " + takeError());
        int cc = cuda_device_compute_capability(deviceID);
        name_ = cuda_device_name(deviceID);
        memory_ = cuda_device_memory(deviceID);
        ccMajor_ = cc / 10;
        ccMinor_ = cc % 10;
    }
    Device(const Device &) = delete;
    Device &operator=(const Device &) = delete;
    ~Device() { Release(); }
    void Release() {  // cuda_bridge.go:470-478
        std::lock_guard<std::mutex> lk(mu_);
        if (ptr_) { cuda_release_device(ptr_); ptr_ = nullptr; }
    }
    int ID() const { return id_; }
    const std::string &Name() const { return name_; }
    uint64_t MemoryBytes() const { return memory_; }
    int MemoryMB() const { return (int)(memory_ / (1024 * 1024)); }
    std::pair<int, int> ComputeCapability() const { return {ccMajor_, ccMinor_}; }

    std::unique_ptr<Buffer> NewBuffer(const std::vector<float> &data, MemoryType memType = MemoryDevice) {  // cuda_bridge.go:506-532
        if (data.empty()) throw CudaError("cuda: cannot create empty buffer");
        std::lock_guard<std::mutex> lk(mu_);
        CudaBuffer *p = cuda_create_buffer(ptr_, const_cast<float *>(data.data()), data.size(), (int)memType);
        if (!p) throw ErrBufferCreation("cuda: failed to create buffer: " + takeError());
        return std::unique_ptr<Buffer>(new Buffer(p, data.size() * 4, this));
    }
    std::unique_ptr<Buffer> NewEmptyBuffer(uint64_t count, MemoryType memType = MemoryDevice) {  // cuda_bridge.go:535-557
        std::lock_guard<std::mutex> lk(mu_);
        CudaBuffer *p = cuda_create_buffer(ptr_, nullptr, (size_t)count, (int)memType);
        if (!p) throw ErrBufferCreation("cuda: failed to create buffer: " + takeError());
        return std::unique_ptr<Buffer>(new Buffer(p, count * 4, this));
    }
    void NormalizeVectors(Buffer &vectors, uint32_t n, uint32_t dimensions) {  // cuda_bridge.go:587-598
        std::lock_guard<std::mutex> lk(mu_);
        if (cuda_normalize_vectors(ptr_, vectors.ptr_, n, dimensions) != 0)
            throw ErrKernelExecution("cuda: kernel execution failed: " + takeError());
        // the buffer was rewritten: the 16-bit image of a cached index must follow
        if (vectors.index_ && nk_index_refresh_shadow(vectors.index_) != 0)
            throw ErrKernelExecution(std::string("cuda: kernel execution failed: ") + nk_last_error());
    }
    void CosineSimilarity(Buffer &embeddings, Buffer &query, Buffer &scores, uint32_t n, uint32_t dimensions, bool normalized) {  // :600-618
        std::lock_guard<std::mutex> lk(mu_);
        if (cuda_cosine_similarity(ptr_, embeddings.ptr_, query.ptr_, scores.ptr_, n, dimensions, normalized ? 1 : 0) != 0)
            throw ErrKernelExecution("cuda: kernel execution failed: " + takeError());
    }
    std::pair<std::vector<uint32_t>, std::vector<float>> TopK(Buffer &scores, uint32_t n, uint32_t k) {  // cuda_bridge.go:620-640
        std::lock_guard<std::mutex> lk(mu_);
        std::vector<uint32_t> idx(k);
        std::vector<float> top(k);
        if (cuda_topk(ptr_, scores.ptr_, idx.data(), top.data(), n, k) != 0)
            throw ErrKernelExecution("cuda: kernel execution failed: " + takeError());
        return {idx, top};
    }
    // cuda_bridge.go:643-686 — ONE fused call instead of NewBuffer + NewEmptyBuffer + CosineSimilarity + TopK.
    // normalized=true keeps the reference's meaning (raw dot of pre-normalised rows); false = true cosine.
    std::vector<SearchResult> Search(Buffer &embeddings, const std::vector<float> &query, uint32_t n, uint32_t dimensions, int k,
                                     bool normalized) {
        auto all = SearchBatch(embeddings, query, 1, n, dimensions, k, normalized ? Metric::Dot : Metric::Cosine);
        return all.empty() ? std::vector<SearchResult>{} : all[0];
    }
    // Q queries (row-major [Q x dimensions]) in one fused launch.
    std::vector<std::vector<SearchResult>> SearchBatch(Buffer &embeddings, const std::vector<float> &queries, uint32_t Q, uint32_t n,
                                                       uint32_t dimensions, int k, Metric metric) {
        if (k <= 0) return {};           // cuda_bridge.go:644-646
        if (k > (int)n) k = (int)n;      // cuda_bridge.go:647-649
        std::lock_guard<std::mutex> lk(mu_);
        if ((uint64_t)n * dimensions * 4 > embeddings.size_) throw ErrInvalidBuffer("cuda: invalid buffer");
        if (!embeddings.index_ || embeddings.index_n_ != n || embeddings.index_dim_ != dimensions || embeddings.index_metric_ != (int)metric) {
            if (embeddings.index_) nk_index_release(embeddings.index_);
            int dev = id_;
            embeddings.index_ = nk_index_create(&dev, 1, dimensions, NK_DTYPE_F32, (int)metric);
            // rows are final by the time they are searched (syncToCUDA = NewBuffer + NormalizeVectors, gpu.go:2073-2118):
            // refresh_shadow gives the attached (caller-owned) rows the fast BF16 filter path
            if (!embeddings.index_ || nk_index_attach_device_rows(embeddings.index_, cuda_buffer_data(embeddings.ptr_), n) != 0 ||
                nk_index_refresh_shadow(embeddings.index_) != 0)
                throw ErrKernelExecution(std::string("cuda: kernel execution failed: ") + nk_last_error());
            embeddings.index_n_ = n; embeddings.index_dim_ = dimensions; embeddings.index_metric_ = (int)metric;
        }
        std::vector<uint32_t> idx((size_t)Q * k);
        std::vector<float> sc((size_t)Q * k);
        int got = nk_search(embeddings.index_, queries.data(), Q, (uint32_t)k, idx.data(), sc.data());
        if (got < 0) throw ErrKernelExecution(std::string("cuda: kernel execution failed: ") + nk_last_error());
        std::vector<std::vector<SearchResult>> out(Q);
        for (uint32_t q = 0; q < Q; ++q)
            for (int i = 0; i < got; ++i) out[q].push_back({idx[(size_t)q * k + i], sc[(size_t)q * k + i]});
        return out;
    }

private:
    CudaDevice *ptr_ = nullptr;
    int id_;
    std::string name_;
    uint64_t memory_ = 0;
    int ccMajor_ = 0, ccMinor_ = 0;
    std::mutex mu_;
};

inline std::unique_ptr<Device> NewDevice(int deviceID) { return std::unique_ptr<Device>(new Device(deviceID)); }  // cuda_bridge.go:445

}  // namespace cuda

namespace gpu {

struct ErrInvalidDimensions : std::invalid_argument { ErrInvalidDimensions() : std::invalid_argument("gpu: invalid dimensions") {} };

struct SearchResult {  // gpu.go SearchResult{ID, Score, Distance}
    std::string ID;
    float Score;
    float Distance;
};

// gpu.EmbeddingIndex (gpu.go:1224-2454) with the device corpus kept in step incrementally (no re-upload on Add/Remove,
// rows stay raw, cosine computed in-kernel).  Node-id strings live on the host only (gpu.go:1228-1231).
class EmbeddingIndex {
public:
    EmbeddingIndex(int dimensions, cuda::Metric metric = cuda::Metric::Cosine, std::vector<int> devices = {0}) : dims_(dimensions) {
        ix_ = nk_index_create(devices.data(), (int)devices.size(), (uint32_t)dimensions, NK_DTYPE_F32, (int)metric);
        if (!ix_) throw cuda::ErrDeviceCreation(std::string("gpu: ") + nk_last_error());
    }
    EmbeddingIndex(const EmbeddingIndex &) = delete;
    EmbeddingIndex &operator=(const EmbeddingIndex &) = delete;
    ~EmbeddingIndex() { Release(); }
    void Release() { if (ix_) { nk_index_release(ix_); ix_ = nullptr; } }

    void Add(const std::string &nodeID, const std::vector<float> &embedding) {  // gpu.go:1378-1403
        if ((int)embedding.size() != dims_) throw ErrInvalidDimensions();
        std::lock_guard<std::mutex> lk(mu_);
        gpuSynced_ = false;
        auto it = idToIndex_.find(nodeID);
        if (it != idToIndex_.end()) check(nk_index_update_row(ix_, it->second, embedding.data()));
        else {
            check(nk_index_append(ix_, embedding.data(), 1));
            idToIndex_[nodeID] = nodeIDs_.size();
            nodeIDs_.push_back(nodeID);
        }
    }
    void AddBatch(const std::vector<std::string> &nodeIDs, const std::vector<std::vector<float>> &embeddings) {  // gpu.go:1406-1434
        if (nodeIDs.size() != embeddings.size()) throw std::invalid_argument("gpu: nodeIDs and embeddings length mismatch");
        for (size_t i = 0; i < nodeIDs.size(); ++i) Add(nodeIDs[i], embeddings[i]);
    }
    bool Remove(const std::string &nodeID) {  // gpu.go:1437-1471 (swap with last)
        std::lock_guard<std::mutex> lk(mu_);
        auto it = idToIndex_.find(nodeID);
        if (it == idToIndex_.end()) return false;
        const uint64_t idx = it->second, last = nodeIDs_.size() - 1;
        gpuSynced_ = false;
        check(nk_index_remove_swap(ix_, idx));
        if (idx != last) {
            nodeIDs_[idx] = nodeIDs_[last];
            idToIndex_[nodeIDs_[idx]] = idx;
        }
        nodeIDs_.pop_back();
        idToIndex_.erase(nodeID);
        return true;
    }
    std::vector<SearchResult> Search(const std::vector<float> &query, int k) {  // gpu.go:1532-1550
        if ((int)query.size() != dims_) throw ErrInvalidDimensions();
        auto r = SearchBatch(query, 1, k);
        return r.empty() ? std::vector<SearchResult>{} : r[0];
    }
    std::vector<std::vector<SearchResult>> SearchBatch(const std::vector<float> &queries, uint32_t Q, int k) {
        if (queries.size() != (size_t)Q * dims_) throw ErrInvalidDimensions();
        std::lock_guard<std::mutex> lk(mu_);
        if (nodeIDs_.empty() || k <= 0) return {};  // gpu.go:1540-1542
        std::vector<uint32_t> idx((size_t)Q * k);
        std::vector<float> sc((size_t)Q * k);
        int got = nk_search(ix_, queries.data(), Q, (uint32_t)k, idx.data(), sc.data());
        if (got < 0) throw cuda::ErrKernelExecution(std::string("gpu: ") + nk_last_error());
        std::vector<std::vector<SearchResult>> out(Q);
        for (uint32_t q = 0; q < Q; ++q)
            for (int i = 0; i < got; ++i) {
                uint32_t r = idx[(size_t)q * k + i];
                float s = sc[(size_t)q * k + i];
                if (r < nodeIDs_.size()) out[q].push_back({nodeIDs_[r], s, 1.0f - s});  // gpu.go:1679-1688
            }
        return out;
    }
    std::vector<SearchResult> ScoreSubset(const std::vector<float> &query, const std::vector<std::string> &ids) {  // gpu.go:1552-1616
        if ((int)query.size() != dims_) throw ErrInvalidDimensions();
        std::lock_guard<std::mutex> lk(mu_);
        std::vector<uint32_t> rows;
        for (auto &id : ids) {
            auto it = idToIndex_.find(id);
            if (it != idToIndex_.end()) rows.push_back((uint32_t)it->second);  // missing ids are ignored
        }
        if (rows.empty()) return {};
        std::vector<uint32_t> idx(rows.size());
        std::vector<float> sc(rows.size());
        int got = nk_score_subset(ix_, query.data(), rows.data(), (uint32_t)rows.size(), (uint32_t)rows.size(), idx.data(), sc.data());
        if (got < 0) throw cuda::ErrKernelExecution(std::string("gpu: ") + nk_last_error());
        std::vector<SearchResult> out;
        for (int i = 0; i < got; ++i) out.push_back({nodeIDs_[idx[i]], sc[i], 1.0f - sc[i]});
        return out;
    }
    // gpu.go:2025: nothing to copy (the device rows are always current); the observable flag of gpu_test.go:1403-1480 is
    // kept: false after Add / Remove / Deserialize, true after SyncToGPU
    void SyncToGPU() { std::lock_guard<std::mutex> lk(mu_); gpuSynced_ = true; }
    bool IsGPUSynced() const { return gpuSynced_; }
    int Count() const { return (int)nodeIDs_.size(); }                                   // gpu.go:2225
    bool Has(const std::string &id) const { return idToIndex_.count(id) != 0; }          // gpu.go:2279
    bool Get(const std::string &id, std::vector<float> *out) {                           // gpu.go:2287
        std::lock_guard<std::mutex> lk(mu_);
        auto it = idToIndex_.find(id);
        if (it == idToIndex_.end()) return false;
        out->resize(dims_);
        check(nk_index_read_rows(ix_, it->second, 1, out->data()));
        return true;
    }
    // gpu.go:2373-2411: LE [dims u32][count u32][len-prefixed ids...][float32 vectors...]
    std::vector<uint8_t> Serialize() {
        std::lock_guard<std::mutex> lk(mu_);
        std::vector<uint8_t> buf;
        auto put32 = [&buf](uint32_t v) { for (int i = 0; i < 4; ++i) buf.push_back((uint8_t)(v >> (8 * i))); };
        put32((uint32_t)dims_);
        put32((uint32_t)nodeIDs_.size());
        for (auto &id : nodeIDs_) { put32((uint32_t)id.size()); buf.insert(buf.end(), id.begin(), id.end()); }
        size_t off = buf.size(), n = nodeIDs_.size();
        buf.resize(off + n * dims_ * 4);
        if (n) check(nk_index_read_rows(ix_, 0, n, buf.data() + off));
        return buf;
    }
    void Deserialize(const std::vector<uint8_t> &data) {  // gpu.go:2414-2454
        if (data.size() < 8) throw std::invalid_argument("gpu: invalid serialized data");
        auto get32 = [&data](size_t o) { return (uint32_t)data[o] | (uint32_t)data[o + 1] << 8 | (uint32_t)data[o + 2] << 16 | (uint32_t)data[o + 3] << 24; };
        uint32_t dims = get32(0), count = get32(4);
        if ((int)dims != dims_) throw ErrInvalidDimensions();
        size_t off = 8;
        std::vector<std::string> ids(count);
        for (uint32_t i = 0; i < count; ++i) {
            uint32_t len = get32(off);
            off += 4;
            ids[i].assign((const char *)&data[off], len);
            off += len;
        }
        std::lock_guard<std::mutex> lk(mu_);
        check(nk_index_upload(ix_, data.data() + off, count));  // the blob streams straight into device memory
        gpuSynced_ = false;
        nodeIDs_ = ids;
        idToIndex_.clear();
        for (uint32_t i = 0; i < count; ++i) idToIndex_[nodeIDs_[i]] = i;
    }

private:
    void check(int rc) { if (rc != 0) throw cuda::ErrKernelExecution(std::string("gpu: ") + nk_last_error()); }
    int dims_;
    NkIndex *ix_ = nullptr;
    std::vector<std::string> nodeIDs_;
    std::unordered_map<std::string, uint64_t> idToIndex_;
    bool gpuSynced_ = false;
    std::mutex mu_;
};

}  // namespace gpu
}  // namespace nornic
