// Device ZSTD decoder: host-side frame walk, scratch layout and launch sequence.  See vl_zstd.cuh for the kernels.
#include "vl_zstd.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <thread>

#include "vl_engine.h"
#include "vl_zstd.cuh"
#include "vl_zstd_job.h"

namespace vl {

using namespace zs;

struct ZstdDev {
    DevBuf frames, blocks, bstate, huf_tab, fse_tab, huf_state, fse_state, predef, lits, seqs, frame_err, status, lists;
    bool predef_ready = false;
    cudaStream_t xstream = nullptr;          // second stream: k_seq_resolve + k_execute of group g run beside the entropy kernels of group g + 1
    std::vector<cudaEvent_t> events;         // grow-only pool (two per launch group: entropy done, execute done)
    cudaEvent_t event(size_t i) {
        while (events.size() <= i) { cudaEvent_t e; VL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); events.push_back(e); }
        return events[i];
    }
    void* hpin = nullptr; size_t hpin_cap = 0;   // page-locked staging of the descriptor tables (read by k_pull_words over PCIe)
    void* ensure_hpin(size_t n) {
        if (n <= hpin_cap) return hpin;
        if (hpin) cudaFreeHost(hpin);
        hpin = nullptr; hpin_cap = 0;
        VL_CUDA(cudaMallocHost(&hpin, n + n / 4 + 4096)); hpin_cap = n + n / 4 + 4096;
        return hpin;
    }
    void release() {
        DevBuf* all[] = {&frames, &blocks, &bstate, &huf_tab, &fse_tab, &huf_state, &fse_state, &predef, &lits, &seqs, &frame_err, &status, &lists};
        for (DevBuf* b : all) b->release();
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
        if (xstream) cudaStreamDestroy(xstream);
        xstream = nullptr;
        if (hpin) cudaFreeHost(hpin);
        hpin = nullptr; hpin_cap = 0;
    }
};

// Descriptor tables reach the device through a kernel that reads page-locked host memory, not through cudaMemcpyAsync: the host->device
// copy engine is busy (and its queue full) with the compressed payload of the same upload, and a copy queued behind it would hold the
// decoder back until the whole payload has landed.
static __global__ void k_pull_words(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
void zstd_dev_free(ZstdDev* d) { if (d) { d->release(); delete d; } }

namespace {
inline unsigned cdiv_u(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }
}  // namespace

struct ZstdJob::Impl : ZstdJobImpl {};

ZstdJob::ZstdJob() : m(new Impl) {}
ZstdJob::~ZstdJob() { delete m; }
bool ZstdJob::empty() const { return m->frames.empty(); }
uint64_t ZstdJob::frames() const { return m->frames.size(); }
uint64_t ZstdJob::blocks() const { return m->blocks.size(); }
uint64_t ZstdJob::compressed_blocks() const { return m->n_compressed; }
uint64_t ZstdJob::sequences() const { return m->n_seqs; }
uint64_t ZstdJob::groups() const { return m->groups.size(); }
void ZstdJob::set_dst(uint32_t id, uint64_t arena_off) { m->frames[id].dst = arena_off; }
void ZstdJob::set_group_hook(std::function<void(uint64_t)> f) { m->group_hook = std::move(f); }

void ZstdJob::add_frame(const uint8_t* f, size_t n, uint64_t zoff, uint64_t* regen, uint32_t* id) {
    ZFrame fr{};
    zwalk::parse_frame_into(m->blocks, (uint32_t)m->frames.size(), f, n, zoff, fr);
    *regen = fr.fcs; *id = (uint32_t)m->frames.size();
    m->commit_frame(fr);
}

size_t ZstdJob::add_bytes_block(const uint8_t* p, size_t n, uint64_t zoff, uint64_t* regen, uint32_t* id) {
    ZFrame fr{};
    const size_t used = zwalk::parse_bytes_block_into(m->blocks, (uint32_t)m->frames.size(), p, n, zoff, fr);
    *regen = fr.fcs; *id = (uint32_t)m->frames.size();
    m->commit_frame(fr);
    return used;
}

void ZstdJob::add_values_blocks(const ZValuesBlock* v, size_t n, int threads, ZValuesInfo* info, size_t* bad, std::string* msg) {
    *bad = SIZE_MAX;
    if (!m->frames.empty()) throw BadInput("BUG: add_values_blocks needs an empty job");
    if (threads > 0) { m->walk_values_blocks(v, n, threads, info, bad, msg); return; }
    // threads == 0: the block-by-block walk the device decoder was first verified with, kept as the reference the threaded walk is compared
    // against (vlscan_zstd_walk_digest).  The tapered group limits need the batch's total number of sequences: a dry walk counts them.
    uint64_t total_seqs = 0;
    if (!group_scale_env()) {
        ZstdJob dry;
        dry.m->limit_scale = 1u << 20;
        for (size_t i = 0; i < n; i++) {
            try {
                uint64_t r = 0; uint32_t id = 0;
                const size_t c1 = dry.add_bytes_block(v[i].p, v[i].n, v[i].zoff, &r, &id);
                dry.add_bytes_block(v[i].p + c1, v[i].n - c1, v[i].zoff + c1, &r, &id);
            } catch (const BadInput&) { break; }   // reported by the walk below
        }
        total_seqs = dry.m->n_seqs;
    }
    auto taper = [&] { if (!group_scale_env()) m->limit_scale = (m->groups.empty() || m->n_seqs > total_seqs - total_seqs / 12) ? 1 : 4; };
    for (size_t i = 0; i < n; i++) {
        try {
            uint32_t f1 = 0, f2 = 0;
            taper();
            const size_t c1 = add_bytes_block(v[i].p, v[i].n, v[i].zoff, &info[i].lens_len, &f1);
            taper();
            const size_t c2 = add_bytes_block(v[i].p + c1, v[i].n - c1, v[i].zoff + c1, &info[i].data_len, &f2);
            if (c1 + c2 != v[i].n) throw BadInput("unexpected non-empty tail after reading bytes block with strings");
        } catch (const BadInput& e) { *bad = i; *msg = e.msg; return; }
    }
    if (!group_scale_env()) m->limit_scale = 1;
}

void ZstdJob::prepare() { m->prepare(); }

void ZstdJob::digest(uint64_t out[4]) const { m->digest(out); }

void ZstdJob::run(vlscan_ctx* ctx, const uint8_t* zsrc, uint8_t* arena) {
    Impl& J = *m;
    if (J.frames.empty()) return;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_prep = now();
    J.prepare();
    const double t_lists = now();
    const std::vector<uint32_t>& lists = J.lists;
    if (!ctx->zdev) ctx->zdev = new ZstdDev;
    ZstdDev& D = *ctx->zdev;
    cudaStream_t st = ctx->stream;
    D.frames.ensure(J.frames.size() * sizeof(ZFrame) + 16); D.blocks.ensure(J.blocks.size() * sizeof(ZBlock) + 16); D.bstate.ensure(J.blocks.size() * sizeof(ZBlockState));
    D.frame_err.ensure(J.frames.size() * 4); D.status.ensure(16); D.lists.ensure(lists.size() * 4 + 16);
    D.huf_tab.ensure(std::max<size_t>((size_t)J.max_huf * Z_HUF_TABLE * 2, 16)); D.huf_state.ensure(std::max<size_t>((size_t)J.max_huf * sizeof(ZSlotState), 16));
    D.fse_tab.ensure(std::max<size_t>((size_t)J.max_fse * Z_FSE_SLOT_BYTES, 16)); D.fse_state.ensure(std::max<size_t>((size_t)J.max_fse * sizeof(ZSlotState), 16));
    // literals and sequence records are produced by the entropy kernels of a group and consumed by its execute kernel on the other stream:
    // two copies, alternating by group, so that the entropy kernels of group g + 1 can run while group g is being executed
    const size_t lits_stride = (J.max_lits + 64 + 255) / 256 * 256, seqs_stride = std::max<size_t>(J.max_seqs, 1);
    D.lits.ensure(2 * lits_stride); D.seqs.ensure(2 * seqs_stride * 16);
    if (!D.xstream) VL_CUDA(cudaStreamCreateWithFlags(&D.xstream, cudaStreamNonBlocking));
    if (!D.predef_ready) {
        D.predef.ensure(Z_FSE_SLOT_BYTES);
        k_zstd_predef<<<1, 32, 0, st>>>(D.predef.as<uint8_t>());
        VL_CUDA(cudaFuncSetAttribute(k_seq_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Z_SEQ_CTA_LANES * (Z_FSE_SLOT_BYTES + Z_LINEBUF))));
        VL_CUDA(cudaFuncSetAttribute(k_huf_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(Z_HUF_CTA_BLOCKS * Z_HUF_TABLE * 2)));
        ctx->launches++; VL_CUDA(cudaGetLastError());
        D.predef_ready = true;
    }
    {
        auto up16 = [](size_t n) { return (n + 15) / 16 * 16; };
        const size_t nf = up16(J.frames.size() * sizeof(ZFrame)), nb = up16(J.blocks.size() * sizeof(ZBlock)), nl = up16(lists.size() * 4);
        uint8_t* h = (uint8_t*)D.ensure_hpin(nf + nb + nl + 64);
        const double t_cp = now();
        J.spread_copy(h, J.frames.data(), J.frames.size() * sizeof(ZFrame));
        J.spread_copy(h + nf, J.blocks.data(), J.blocks.size() * sizeof(ZBlock));
        if (!lists.empty()) J.spread_copy(h + nf + nb, lists.data(), lists.size() * 4);
        if (getenv("VLSCAN_DEBUG_TIMING")) fprintf(stderr, "[vlscan zstd] host: work lists of %zu launch groups %.1f ms, %.1f MB of tables to pinned memory %.1f ms (%d threads)\n",
                                                   J.groups.size(), 1e3 * (t_lists - t_prep), (nf + nb + nl) / 1e6, 1e3 * (now() - t_cp), std::max(J.threads, 1));
        k_pull_words<<<296, 256, 0, st>>>(D.frames.as<uint4>(), (const uint4*)h, nf / 16); ctx->launches++;
        k_pull_words<<<296, 256, 0, st>>>(D.blocks.as<uint4>(), (const uint4*)(h + nf), nb / 16); ctx->launches++;
        if (nl) { k_pull_words<<<296, 256, 0, st>>>(D.lists.as<uint4>(), (const uint4*)(h + nf + nb), nl / 16); ctx->launches++; }
        VL_CUDA(cudaGetLastError());
    }
    VL_CUDA(cudaMemsetAsync(D.frame_err.p, 0, J.frames.size() * 4, st));
    VL_CUDA(cudaMemsetAsync(D.status.p, 0, 16, st));
    VL_CUDA(cudaMemsetAsync(D.bstate.p, 0, J.blocks.size() * sizeof(ZBlockState), st));
    ZView V{};
    V.src = zsrc; V.arena = arena; V.frames = D.frames.as<ZFrame>(); V.blocks = D.blocks.as<ZBlock>(); V.bstate = D.bstate.as<ZBlockState>();
    V.huf_tab = D.huf_tab.as<uint16_t>(); V.fse_tab = D.fse_tab.as<uint8_t>(); V.huf_state = D.huf_state.as<ZSlotState>(); V.fse_state = D.fse_state.as<ZSlotState>();
    V.predef = D.predef.as<uint8_t>(); V.lits = D.lits.as<uint8_t>(); V.seqs = D.seqs.as<uint4>(); V.frame_err = D.frame_err.as<unsigned int>();
    V.status = D.status.as<unsigned long long>();
    const uint32_t* L = D.lists.as<uint32_t>();
    // VLSCAN_DEBUG_TIMING: device time per phase (events around every launch; summed over the groups)
    const bool dbg = getenv("VLSCAN_DEBUG_TIMING") != nullptr;
    static const char* phase_name[6] = {"huf_build", "huf_decode", "fse_build", "seq_decode", "seq_resolve", "execute"};
    std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> marks;
    int phase = 0;
    // VLSCAN_ZSTD_OVERLAP=0 (tuning only) keeps resolve + execute on the ctx stream
    static const bool overlap = [] { const char* e = getenv("VLSCAN_ZSTD_OVERLAP"); return !e || atoi(e) != 0; }();
    cudaStream_t xs = overlap ? D.xstream : st;
    cudaStream_t cur_stream = st;
    auto begin = [&](int p) { phase = p; if (dbg) { cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, cur_stream); marks.push_back({p, {a, b}}); } };
    auto launched = [&] { ctx->launches++; VL_CUDA(cudaGetLastError()); if (dbg) cudaEventRecord(marks.back().second.second, cur_stream); };
    // everything enqueued so far on the ctx stream (tables, memsets) precedes the first execute
    VL_CUDA(cudaEventRecord(D.event(0), st)); VL_CUDA(cudaStreamWaitEvent(xs, D.event(0), 0));
    size_t gi = 0;
    for (const Group& g : J.groups) {
        uint32_t nh = g.huf_hi - g.huf_lo, nl = g.lit_hi - g.lit_lo, ns = g.seq_hi - g.seq_lo, nf = g.frame_hi - g.frame_lo;
        if (J.group_hook) {
            uint64_t need = 0;
            for (uint32_t i = J.frames[g.frame_lo].blk_lo; i < J.frames[g.frame_hi - 1].blk_hi; i++) need = std::max<uint64_t>(need, J.blocks[i].src + (J.blocks[i].type == ZB_RLE ? 1 : J.blocks[i].size));
            J.group_hook(need);
        }
        ZView W = V;
        W.lits = V.lits + (gi & 1) * lits_stride; W.seqs = V.seqs + (gi & 1) * seqs_stride;
        cudaEvent_t ev_entropy = D.event(1 + 2 * gi), ev_exec = D.event(2 + 2 * gi);
        cur_stream = st;
        if (gi >= 2) VL_CUDA(cudaStreamWaitEvent(st, D.event(2 + 2 * (gi - 2)), 0));   // the scratch copy this group writes was read by group gi - 2
        if (nh) { begin(0); k_huf_build<<<cdiv_u(nh, 4), 128, 0, st>>>(W, L + g.huf_lo, nh); launched(); }
        if (nl) { begin(1); k_huf_decode<<<cdiv_u(nl, Z_HUF_CTA_BLOCKS), Z_HUF_CTA_BLOCKS * 4, Z_HUF_CTA_BLOCKS * Z_HUF_TABLE * 2, st>>>(W, L + g.lit_lo, nl); launched(); }
        if (ns) {
            begin(2); k_fse_build<<<cdiv_u(ns, 64), 64, 0, st>>>(W, L + g.seq_lo, ns); launched();
            begin(3); k_seq_decode<<<cdiv_u(ns, Z_SEQ_CTA_LANES), 64, Z_SEQ_CTA_LANES * (Z_FSE_SLOT_BYTES + Z_LINEBUF), st>>>(W, L + g.seq_lo, ns); launched();
        }
        VL_CUDA(cudaEventRecord(ev_entropy, st));
        VL_CUDA(cudaStreamWaitEvent(xs, ev_entropy, 0));
        cur_stream = xs;
        begin(4); k_seq_resolve<<<cdiv_u(nf, 64), 64, 0, xs>>>(W, g.frame_lo, nf); launched();
        begin(5); k_execute<<<cdiv_u((uint64_t)nf * 32, Z_EXEC_WARPS * 32), Z_EXEC_WARPS * 32, 0, xs>>>(W, L + g.ord_lo, nf); launched();
        VL_CUDA(cudaEventRecord(ev_exec, xs));
        gi++;
    }
    if (gi) VL_CUDA(cudaStreamWaitEvent(st, D.event(2 + 2 * (gi - 1)), 0));   // the ctx stream continues behind the last execute
    cur_stream = st;
    if (dbg) {
        VL_CUDA(cudaStreamSynchronize(xs)); VL_CUDA(cudaStreamSynchronize(st));
        float tot[6] = {0, 0, 0, 0, 0, 0};
        for (auto& mk : marks) { float ms = 0; cudaEventElapsedTime(&ms, mk.second.first, mk.second.second); tot[mk.first] += ms; cudaEventDestroy(mk.second.first); cudaEventDestroy(mk.second.second); }
        fprintf(stderr, "[vlscan zstd] %zu groups:", J.groups.size());
        for (int p = 0; p < 6; p++) fprintf(stderr, " %s %.2f ms", phase_name[p], tot[p]);
        fprintf(stderr, "\n");
    }
    (void)phase;
    J.ran = true;
}

void ZstdJob::check(vlscan_ctx* ctx) {
    if (!m->ran) return;
    unsigned long long stt[2] = {0, 0};
    VL_CUDA(cudaMemcpyAsync(stt, ctx->zdev->status.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
    VL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (!stt[0]) return;
    static const char* what[] = {"", "bad Huffman tree description", "bad Huffman stream", "bad FSE table description", "bad sequences bitstream", "match offset beyond the regenerated data",
                                 "regenerated size differs from the frame content size", "bad literals section"};
    throw BadInput(std::string("cannot decompress block: ") + (stt[0] < 8 ? what[stt[0]] : "corrupted frame") + " (frame " + std::to_string(stt[1] - 1) + " of the batch)");
}

}  // namespace vl
