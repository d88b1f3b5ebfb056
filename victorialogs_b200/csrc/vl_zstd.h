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

// one on-disk values block: bytesBlock(lens) ++ bytesBlock(data) (stringsBlockUnmarshaler.unmarshal, lib/logstorage/encoding.go:83-108);
// its first byte will sit at offset `zoff` of the compressed staging buffer
struct ZValuesBlock { const uint8_t* p; size_t n; uint64_t zoff; };
struct ZValuesInfo { uint64_t lens_len, data_len; };   // regenerated lengths of the two bytes blocks

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
    // Registers n on-disk values blocks of an empty job: block i becomes frames 2i (lens) and 2i+1 (data).  threads >= 1 walks them on that
    // many host threads, 0 walks them one by one through add_bytes_block; the job comes out the same either way.  Malformed input does not
    // throw: *bad gets the index of the first bad block (else SIZE_MAX) and *msg what a sequential walk would have thrown there, so that
    // the caller can raise it where its own block-by-block validation reaches that block.
    void add_values_blocks(const ZValuesBlock* v, size_t n, int threads, ZValuesInfo* info, size_t* bad, std::string* msg);
    // host half of run(): closes the last launch group and builds the work lists (idempotent)
    void prepare();
    // digest of everything run() hands to the device: frames, blocks, launch groups + scratch sizes, work lists
    void digest(uint64_t out[4]) const;
    void set_dst(uint32_t id, uint64_t arena_off);
    bool empty() const;
    uint64_t frames() const;
    uint64_t blocks() const;
    uint64_t compressed_blocks() const;
    uint64_t sequences() const;
    uint64_t groups() const;    // launch groups (after prepare)
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
