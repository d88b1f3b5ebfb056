// Host side of the device ZSTD decoder (kernels in vl_zstd.cuh): walks bytes-block containers, frame headers and block headers
// (the only parts of a frame the host ever reads), lays out scratch, and enqueues the decode phases on the ctx stream.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <functional>
#include <string>
#include <vector>

struct vlscan_ctx;

namespace vl {

struct ZstdDev;                 // per-ctx device scratch of the decoder (grow-only)
void zstd_dev_free(ZstdDev* d);

struct ZstdTimings { float total_ms = 0; };

class ZstdJob {
public:
    ZstdJob();
    ~ZstdJob();
    ZstdJob(const ZstdJob&) = delete;
    ZstdJob& operator=(const ZstdJob&) = delete;
    // unmarshalBytesBlock (lib/logstorage/encoding.go:372-426) without the decompression: registers the bytes block that starts at
    // host address p (n bytes available) and whose first byte will sit at offset `zoff` of the compressed staging buffer.
    // Returns the bytes consumed; *regen = regenerated length; *id = handle for set_dst.  Throws BadInput on malformed containers,
    // frame headers and block headers.
    size_t add_bytes_block(const uint8_t* p, size_t n, uint64_t zoff, uint64_t* regen, uint32_t* id);
    // a bare ZSTD frame occupying exactly [f, f+n)
    void add_frame(const uint8_t* f, size_t n, uint64_t zoff, uint64_t* regen, uint32_t* id);
    void set_dst(uint32_t id, uint64_t arena_off);
    bool empty() const;
    uint64_t frames() const;
    uint64_t blocks() const;
    uint64_t compressed_blocks() const;
    uint64_t sequences() const;
    // called right before the kernels of a launch group are enqueued, with the end offset (in the compressed staging buffer) of the
    // last byte the group reads: lets the caller make the stream wait for exactly that part of an upload still in flight
    void set_group_hook(std::function<void(uint64_t src_end)> f);
    // enqueue all decode phases on ctx->stream; zsrc / arena are device pointers
    void run(vlscan_ctx* ctx, const uint8_t* zsrc, uint8_t* arena);
    // after the stream was synchronised: throws BadInput("cannot decompress block: ...") if a frame failed on the device
    void check(vlscan_ctx* ctx);
private:
    struct Impl;
    Impl* m;
};

}  // namespace vl
