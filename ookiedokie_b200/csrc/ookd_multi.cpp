// ookd_multi.cpp -- one capture window time-sharded over several GPUs from ONE process (ookd_gpu_multi_* in
// include/ookd_gpu.h; SURVEY 8(b)'s `gpu_ids, n_gpus`, 8(e)-2).
//
// Pure orchestration over the single-GPU C ABI: shard g = consecutive whole multiples of lcm(samples_per_buffer,
// decimation) samples, decoded by handle g on its own host thread (ookd_gpu_decode_shard: FIR halo read from the
// samples in front of the shard, state machine entered from one chunk of warm-up history), then stitched on the host:
// a shard whose predecessor's exit state differs from the entry it assumed re-runs ONLY its state-machine stage
// (ookd_gpu_resolve).  No collective, no peer copies: what crosses GPUs is one 48-byte carry per boundary and the
// message lists.  The reference has no counterpart (it is single threaded, src/ookiedokie.c:238-290); the result is
// the one its loop would print for the whole window.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "ookd_gpu.h"

struct ookd_gpu_multi {
    std::vector<ookd_gpu *> h;
    uint32_t halo = 0, dec = 1, spb = 1;
    uint64_t align = 1;
    uint32_t used = 0;                       // shards of the last decode
    std::vector<ookd_msg> msgs;
    std::vector<uint64_t> edges;
    std::vector<ookd_gpu_result> res;
    std::vector<ookd_sm_carry> exits;
    std::vector<int> status;
    uint32_t resolves = 0;
    char err[256] = {0};
};

namespace {

uint64_t gcd64(uint64_t a, uint64_t b)
{
    while (b) {
        const uint64_t t = a % b;
        a = b;
        b = t;
    }
    return a;
}

int mfail(ookd_gpu_multi *m, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(m->err, sizeof(m->err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace

extern "C" {

int ookd_gpu_multi_create(ookd_gpu_multi **out, const struct ookd_gpu_config *cfg, const int32_t *gpu_ids, uint32_t n_gpus)
{
    if (!out || !cfg || !gpu_ids || n_gpus == 0 || n_gpus > 64) return OOKD_ERR_ARG;
    *out = nullptr;
    ookd_gpu_multi *m = new ookd_gpu_multi();
    for (uint32_t g = 0; g < n_gpus; g++) {
        ookd_gpu_config c = *cfg;
        c.device_id = gpu_ids[g];
        c.sm_warmup = (n_gpus > 1 && cfg->sm) ? 1u : cfg->sm_warmup;     // provisional entries from one chunk of history
        ookd_gpu *h = nullptr;
        const int rc = ookd_gpu_create(&h, &c);
        if (rc != OOKD_OK) {
            ookd_gpu_multi_destroy(m);
            return rc;
        }
        m->h.push_back(h);
    }
    m->halo = ookd_gpu_halo(m->h[0]);
    m->dec = ookd_gpu_total_decimation(m->h[0]);
    m->spb = cfg->samples_per_buffer;
    m->align = (uint64_t) m->spb / gcd64(m->spb, m->dec) * m->dec;
    m->res.resize(n_gpus);
    m->exits.resize(n_gpus);
    m->status.assign(n_gpus, OOKD_OK);
    *out = m;
    return OOKD_OK;
}

void ookd_gpu_multi_destroy(ookd_gpu_multi *m)
{
    if (!m) return;
    for (ookd_gpu *h : m->h) ookd_gpu_destroy(h);
    delete m;
}

uint32_t ookd_gpu_multi_halo(const ookd_gpu_multi *m) { return m ? m->halo : 0; }
uint32_t ookd_gpu_multi_n_gpus(const ookd_gpu_multi *m) { return m ? (uint32_t) m->h.size() : 0; }
ookd_gpu *ookd_gpu_multi_handle(ookd_gpu_multi *m, uint32_t g) { return (m && g < m->h.size()) ? m->h[g] : nullptr; }
const char *ookd_gpu_multi_last_error(const ookd_gpu_multi *m) { return m ? m->err : "null handle"; }
uint32_t ookd_gpu_multi_shards_used(const ookd_gpu_multi *m) { return m ? m->used : 0; }

int ookd_gpu_multi_shard_range(const ookd_gpu_multi *m, uint64_t first_sample, uint64_t n_samples, uint32_t g,
                               uint64_t *shard_first, uint64_t *shard_n)
{
    if (!m || !shard_first || !shard_n || g >= m->h.size()) return OOKD_ERR_ARG;
    const uint64_t G = m->h.size();
    const uint64_t units = (n_samples + m->align - 1) / m->align;
    const uint64_t per = ((units + G - 1) / G) * m->align;                // samples per shard (last one may be shorter)
    const uint64_t lo = (uint64_t) g * per;
    *shard_first = first_sample + (lo < n_samples ? lo : n_samples);
    *shard_n = lo >= n_samples ? 0 : (n_samples - lo < per ? n_samples - lo : per);
    return OOKD_OK;
}

int ookd_gpu_multi_decode(ookd_gpu_multi *m, const void *iq, int iq_is_device_ptrs, uint64_t first_sample, uint64_t n_samples,
                          int last, const struct ookd_sm_carry *entry, struct ookd_sm_carry *exit_, struct ookd_gpu_result *res)
{
    if (!m || (!iq && n_samples)) return OOKD_ERR_ARG;
    if (first_sample % m->align) return mfail(m, OOKD_ERR_ARG, "first_sample must be a multiple of lcm(spb, decimation)");
    if (!last && (n_samples % m->align)) return mfail(m, OOKD_ERR_ARG, "non-final window length must be a multiple of lcm(spb, decimation)");
    const uint32_t G = (uint32_t) m->h.size();
    std::vector<uint64_t> sf(G), sn(G);
    uint32_t used = 0;
    for (uint32_t g = 0; g < G; g++) {
        ookd_gpu_multi_shard_range(m, first_sample, n_samples, g, &sf[g], &sn[g]);
        if (sn[g] > 0 || g == 0) used = g + 1;
    }
    m->used = used;
    m->resolves = 0;
    const uint64_t halo_avail0 = first_sample < m->halo ? first_sample : m->halo;

    auto worker = [&](uint32_t g) {
        const uint64_t ha = sf[g] < m->halo ? sf[g] : m->halo;
        const int16_t *p;
        if (iq_is_device_ptrs) {
            p = ((const int16_t *const *) iq)[g];                          // already points at sf[g] - ha on GPU g
        } else {
            // host window: iq[0] is sample first_sample - halo_avail0
            p = (const int16_t *) iq + 2 * ((sf[g] - ha) - (first_sample - halo_avail0));
        }
        const bool is_last = last && (g + 1 == used);
        m->status[g] = ookd_gpu_decode_shard(m->h[g], p, iq_is_device_ptrs ? 1 : 0, sf[g], sn[g], is_last ? 1 : 0,
                                             g == 0 ? entry : nullptr, &m->exits[g], &m->res[g]);
    };
    std::vector<std::thread> threads;
    for (uint32_t g = 1; g < used; g++) threads.emplace_back(worker, g);
    worker(0);
    for (auto &t : threads) t.join();
    for (uint32_t g = 0; g < used; g++) {
        if (m->status[g]) return mfail(m, m->status[g], "shard %u: %s", g, ookd_gpu_last_error(m->h[g]));
    }
    // ---- stitch: shard g must have been entered in the state shard g-1 was left in ----
    for (uint32_t g = 1; g < used; g++) {
        if (memcmp(&m->res[g].entry_used, &m->exits[g - 1], sizeof(ookd_sm_carry)) != 0) {
            int rc = ookd_gpu_resolve(m->h[g], &m->exits[g - 1], &m->exits[g], &m->res[g]);
            if (rc == OOKD_ERR_STATE) {
                // the shard's tables cannot take the corrected entry (a long cascade): decode it again, entered explicitly
                const uint64_t ha = sf[g] < m->halo ? sf[g] : m->halo;
                const int16_t *p = iq_is_device_ptrs ? ((const int16_t *const *) iq)[g]
                                                     : (const int16_t *) iq + 2 * ((sf[g] - ha) - (first_sample - halo_avail0));
                rc = ookd_gpu_decode_shard(m->h[g], p, iq_is_device_ptrs ? 1 : 0, sf[g], sn[g], (last && g + 1 == used) ? 1 : 0,
                                           &m->exits[g - 1], &m->exits[g], &m->res[g]);
            }
            if (rc) return mfail(m, rc, "resolve of shard %u: %s", g, ookd_gpu_last_error(m->h[g]));
            m->resolves++;
        }
    }
    m->msgs.clear();
    for (uint32_t g = 0; g < used; g++) {
        if (m->res[g].n_msgs) m->msgs.insert(m->msgs.end(), m->res[g].msgs, m->res[g].msgs + m->res[g].n_msgs);
    }
    if (exit_) *exit_ = m->exits[used - 1];
    if (res) {
        memset(res, 0, sizeof(*res));
        for (uint32_t g = 0; g < used; g++) {
            const ookd_gpu_result &r = m->res[g];
            res->n_in += r.n_in;
            res->n_out += r.n_out;
            res->n_buffers += r.n_buffers;
            res->n_edges += r.n_edges;
            res->gpu_launches += r.gpu_launches;
            res->host_syncs += r.host_syncs;
            res->refined_blocks += r.refined_blocks;
            res->refined_tiles += r.refined_tiles;
            if (r.kernel_ms > res->kernel_ms) res->kernel_ms = r.kernel_ms;      // shards run concurrently
            if (r.fir_ms > res->fir_ms) res->fir_ms = r.fir_ms;
            if (r.screen_ms > res->screen_ms) res->screen_ms = r.screen_ms;
            if (r.sm_rounds > res->sm_rounds) res->sm_rounds = r.sm_rounds;
        }
        res->first_bit = m->res[0].first_bit;
        res->entry_used = m->res[0].entry_used;
        res->n_msgs = m->msgs.size();
        res->msgs = m->msgs.empty() ? nullptr : m->msgs.data();
        res->sm_rounds += m->resolves;
    }
    return OOKD_OK;
}

int ookd_gpu_multi_edges(ookd_gpu_multi *m, const uint64_t **edges, uint64_t *n_edges, uint32_t *first_bit)
{
    if (!m || !edges || !n_edges) return OOKD_ERR_ARG;
    m->edges.clear();
    for (uint32_t g = 0; g < m->used; g++) {
        const uint64_t *e = nullptr;
        uint64_t n = 0;
        uint32_t fb = 0;
        const int rc = ookd_gpu_edges(m->h[g], &e, &n, &fb);
        if (rc) return mfail(m, rc, "edges of shard %u: %s", g, ookd_gpu_last_error(m->h[g]));
        if (g == 0 && first_bit) *first_bit = fb;
        m->edges.insert(m->edges.end(), e, e + n);
    }
    *edges = m->edges.data();
    *n_edges = m->edges.size();
    return OOKD_OK;
}

}  // extern "C"
