// Multi-GPU layout of the clustersearch hot path in the C ABI: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The path shards over QUERY SETS: every rank holds the whole target index (BASELINE's largest config is 4.3 GB of the 180 GB)
// and owns a contiguous, residue-balanced block of query sets (the analogue of DBReader::decomposeDomainByAminoAcid,
// lib/mmseqs/src/commons/DBReader.cpp).  Every stage is independent per query set once the whole target side is visible -- a
// query's prefilter list depends on the query and the FULL target index (the --max-seqs cut, its tie order and the hit-buffer
// overflow are global over targets, QueryMatcher.cpp:43,108-211,280-333), E-values use the total target residues
// (Alignment.cpp:263), combinehits multiplies by the number of target sets (src/util/combinehits.cpp:77,153) -- so there is no
// collective inside the data path.  After the step ONE ncclAllGather (sizes, then padded blocks) leaves every rank with the whole
// job's result lists.  The reference's own multi-node mode splits the target DB instead and merges per-split prefilter lists
// (Prefiltering.cpp:273-377,662-723) with a per-split --max-seqs (N/split + 4 sqrt(N/split), :358-361), which does not give the
// unsplit result; see DESIGN.md section 5 for the measured target-sharded ablation.
//
// NCCL is resolved at run time (dlopen): single-GPU users of libsdgpu.so need no NCCL.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <memory>

#include "sd_internal.h"

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi &nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.err = dlerror() ? dlerror() : "libnccl.so.2 not found"; return; }
        auto sym = [&](const char *n) { void *p = dlsym(api.handle, n); if (!p && api.err.empty()) api.err = std::string("missing symbol ") + n; return p; };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api;
}

int ncclReady() {
    NcclApi &n = nccl();
    if (!n.handle || !n.err.empty()) { sd::setError("NCCL is not available: %s", n.err.c_str()); return SD_ENODEV; }
    return SD_OK;
}

#define SD_NCCL(call)                                                                                    \
    do {                                                                                                 \
        ncclResult_t r__ = (call);                                                                       \
        if (r__ != ncclSuccess) {                                                                        \
            sd::setError("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r__)); \
            return SD_ENODEV;                                                                            \
        }                                                                                                \
    } while (0)

static_assert(sizeof(ncclUniqueId) == SD_NCCL_ID_BYTES, "sdgpu.h: SD_NCCL_ID_BYTES");

inline uint64_t align16(uint64_t x) { return (x + 15) & ~15ull; }

}  // namespace

extern "C" int sd_shard_unique_id(uint8_t *id) {
    if (!id) { sd::setError("sd_shard_unique_id: null argument"); return SD_EINVAL; }
    int rc = ncclReady();
    if (rc) return rc;
    ncclUniqueId u;
    SD_NCCL(nccl().GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return SD_OK;
}

extern "C" int sd_shard_init(sd_context *ctx, int rank, int world, const uint8_t *id) {
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) { sd::setError("sd_shard_init: bad argument"); return SD_EINVAL; }
    int rc = ncclReady();
    if (rc) return rc;
    if (ctx->comm) { sd::setError("sd_shard_init: the context already has a communicator"); return SD_EINVAL; }
    SD_CUDA(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm;
    SD_NCCL(nccl().CommInitRank(&comm, world, u, rank));
    ctx->comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    // one process per GPU of one node: this rank's share of the host threads (SD_HOST_THREADS overrides)
    sd::hostThreadLimit() = getenv("SD_HOST_THREADS") ? (unsigned)std::max(1, atoi(getenv("SD_HOST_THREADS")))
                                                      : std::max(4u, std::thread::hardware_concurrency() / (unsigned)world);
    return SD_OK;
}

extern "C" void sd_shard_finalize(sd_context *ctx) {
    if (!ctx || !ctx->comm) return;
    cudaSetDevice(ctx->device);
    nccl().CommDestroy(static_cast<ncclComm_t>(ctx->comm));
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
}

extern "C" int sd_shard_rank(const sd_context *ctx) { return ctx ? ctx->rank : 0; }
extern "C" int sd_shard_world(const sd_context *ctx) { return ctx ? ctx->world : 1; }

// Contiguous blocks of query sets with about equal residues: first_set[r] .. first_set[r+1] belong to rank r.  Split at the
// prefix-sum quantiles (the closer of the two neighbouring boundaries); every rank gets at least one set while sets remain, so with
// fewer sets than ranks the LAST ranks are empty and rank 0 always has work.
extern "C" int sd_shard_partition(const uint64_t *set_residues, uint32_t n_sets, int world, uint32_t *first_set) {
    if (!set_residues || !first_set || world < 1) { sd::setError("sd_shard_partition: bad argument"); return SD_EINVAL; }
    std::vector<double> pre(n_sets + 1, 0.0);
    for (uint32_t i = 0; i < n_sets; i++) pre[i + 1] = pre[i] + (double)set_residues[i];
    const double total = pre[n_sets];
    first_set[0] = 0;
    for (int r = 1; r < world; r++) {
        const double target = total * r / world;
        uint32_t c = (uint32_t)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
        if (c > 0 && c <= n_sets && std::abs(pre[c - 1] - target) < std::abs(pre[c] - target)) c--;
        c = std::min(c, n_sets);
        const uint32_t loBound = std::min(first_set[r - 1] + 1, n_sets);
        const uint32_t remaining = (uint32_t)(world - r);
        const uint32_t hiBound = std::max(n_sets > remaining ? n_sets - remaining : 0u, loBound);
        first_set[r] = std::max(loBound, std::min(c, hiBound));
    }
    first_set[world] = n_sets;
    return SD_OK;
}

// Contiguous blocks of query GENES with about equal residues (gene_off: n_genes + 1 residue offsets): rank r owns genes
// first_gene[r] .. first_gene[r+1]-1.  The unit of the prefilter and align stages of sd_shard_clustersearch_run.
extern "C" int sd_shard_partition_genes(const uint64_t *gene_off, uint32_t n_genes, int world, uint32_t *first_gene) {
    if (!gene_off || !first_gene || world < 1) { sd::setError("sd_shard_partition_genes: bad argument"); return SD_EINVAL; }
    const uint64_t base = gene_off[0], total = gene_off[n_genes] - base;
    first_gene[0] = 0;
    for (int r = 1; r < world; r++) {
        const uint64_t target = base + (uint64_t)((double)total * r / world);
        uint32_t c = (uint32_t)(std::lower_bound(gene_off, gene_off + n_genes + 1, target) - gene_off);
        c = std::min(c, n_genes);
        first_gene[r] = std::max(first_gene[r - 1], c);
    }
    first_gene[world] = n_genes;
    return SD_OK;
}

// Block layout: [4 x u64 counts][alignments][backtrace pool][clusters][members], every part on a 16-byte boundary.
// ONE exchange: ncclAllGather of the block sizes, then ncclAllGather of the blocks padded to the largest one
// (pinned host -> HBM -> NVLink -> HBM -> pinned host); the merged result has one chunk per rank, in rank (= query set) order.
extern "C" int sd_shard_allgather_result(sd_context *ctx, const sd_search_result *mine, uint32_t query_first, sd_search_result **all,
                                         uint64_t *bytes_gathered) {
    return sd::shardAllgather(ctx, mine, query_first, 1, all, bytes_gathered);
}

// pinSlot, pinSlot + 1: pinned send / receive blocks of this exchange (the merged result is a view of the receive block, so a step
// with two exchanges uses two pairs)
int sd::shardAllgather(sd_context *ctx, const sd_search_result *mine, uint32_t query_first, int pinSlot, sd_search_result **all,
                       uint64_t *bytes_gathered) {
    if (!ctx || !mine || !all) { sd::setError("sd_shard_allgather_result: bad argument"); return SD_EINVAL; }
    *all = nullptr;
    if (!ctx->comm) { sd::setError("sd_shard_allgather_result: call sd_shard_init first"); return SD_EINVAL; }
    SD_CUDA(cudaSetDevice(ctx->device));
    ncclComm_t comm = static_cast<ncclComm_t>(ctx->comm);
    const int world = ctx->world;
    cudaStream_t st = ctx->stream;
    uint64_t na = 0, np = 0, nc = 0, nm = 0;
    sd_search_result_sizes(mine, &na, &np, &nc, &nm);
    const uint64_t oA = 32, oP = align16(oA + na * sizeof(sd_alignment)), oC = align16(oP + np), oM = align16(oC + nc * sizeof(sd_cluster));
    const uint64_t used = align16(oM + nm * sizeof(sd_cluster_member));

    // sizes
    uint64_t *dSizes = static_cast<uint64_t *>(sd::ctxDeviceScratch(ctx, sd::SL_SH_SIZES, (size_t)(world + 1) * sizeof(uint64_t)));
    if (!dSizes) return SD_ENOMEM;
    SD_CUDA(cudaMemcpyAsync(dSizes + world, &used, sizeof used, cudaMemcpyHostToDevice, st));
    SD_NCCL(nccl().AllGather(dSizes + world, dSizes, sizeof(uint64_t), ncclUint8, comm, st));
    std::vector<uint64_t> sizes(world);
    SD_CUDA(cudaMemcpyAsync(sizes.data(), dSizes, (size_t)world * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));
    uint64_t cap = 0;
    for (uint64_t s : sizes) cap = std::max(cap, s);
    cap = (cap + 4095) & ~4095ull;

    // local block in pinned memory (cross references stay block-relative; the query index becomes global)
    char *hSend = static_cast<char *>(sd::ctxPinnedScratch(ctx, pinSlot, cap));
    char *hRecv = static_cast<char *>(sd::ctxPinnedScratch(ctx, pinSlot + 1, cap * (uint64_t)world));
    char *dSend = static_cast<char *>(sd::ctxDeviceScratch(ctx, sd::SL_SH_SEND, cap));
    char *dRecv = static_cast<char *>(sd::ctxDeviceScratch(ctx, sd::SL_SH_RECV, cap * (uint64_t)world));
    if (!hSend || !hRecv || !dSend || !dRecv) return SD_ENOMEM;
    uint64_t head[4] = {na, np, nc, nm};
    memcpy(hSend, head, sizeof head);
    int rc = sd_search_result_copy(mine, reinterpret_cast<sd_alignment *>(hSend + oA), hSend + oP, reinterpret_cast<sd_cluster *>(hSend + oC),
                                   reinterpret_cast<sd_cluster_member *>(hSend + oM));
    if (rc) return rc;
    if (query_first) {
        sd_alignment *a = reinterpret_cast<sd_alignment *>(hSend + oA);
        sd::parallelFor(na, [&](uint64_t i) { a[i].query += query_first; });
    }
    SD_CUDA(cudaMemcpyAsync(dSend, hSend, used, cudaMemcpyHostToDevice, st));
    SD_NCCL(nccl().AllGather(dSend, dRecv, cap, ncclUint8, comm, st));
    for (int r = 0; r < world; r++)      // only the used part of every block comes back to the host
        SD_CUDA(cudaMemcpyAsync(hRecv + (uint64_t)r * cap, dRecv + (uint64_t)r * cap, sizes[r], cudaMemcpyDeviceToHost, st));
    SD_CUDA(cudaStreamSynchronize(st));

    rc = sd_search_result_from_blocks(hRecv, cap, sizes.data(), world, all);
    if (bytes_gathered) { *bytes_gathered = 0; for (uint64_t s : sizes) *bytes_gathered += s; }
    return rc;
}
