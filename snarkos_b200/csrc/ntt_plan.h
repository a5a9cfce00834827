// ntt_plan.h -- pass planning for the multi-pass NTT (plain C++: shared by ntt.cu and the host test shim).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef NTT_MAX_PASSES
#define NTT_MAX_PASSES 4
#endif
#define NTT_MAX_TILE_LOG 12                 // 4096 elements = 128 KB of shared memory

// ---------------------------------------------------------------------------------------------
// pass planning: split log_n into pass lengths, pick the tile width of every pass
// ---------------------------------------------------------------------------------------------
struct NttPlan {
    uint32_t npasses;
    uint32_t log_len[NTT_MAX_PASSES];
    uint32_t log_cw[NTT_MAX_PASSES];
};

// plan_override "a,b,c" (option ntt_plan) fixes the pass lengths; tile_log (option ntt_tile_log, default 11: 2048
// elements = 64 KB -> 3 CTAs per SM) is the tile size used when splitting.
static bool ntt_make_plan(uint32_t log_n, NttPlan* plan, uint32_t tile_log = 11, const char* plan_override = nullptr) {
    memset(plan, 0, sizeof(*plan));
    if (tile_log > NTT_MAX_TILE_LOG) tile_log = NTT_MAX_TILE_LOG;
    if (tile_log < 4) tile_log = 4;
    bool have = false;
    if (const char* e = (plan_override && *plan_override) ? plan_override : nullptr) {
        uint32_t sum = 0, k = 0;
        const char* q = e;
        while (*q && k < NTT_MAX_PASSES) {
            uint32_t v = (uint32_t)strtoul(q, (char**)&q, 10);
            if (v == 0) break;
            plan->log_len[k++] = v;
            sum += v;
            if (*q == ',') q++;
        }
        if (sum == log_n && k >= 1) { plan->npasses = k; have = true; }
    }
    if (!have) {
        if (log_n <= NTT_MAX_TILE_LOG) {
            plan->npasses = 1;
            plan->log_len[0] = log_n;
        } else {
            // each pass keeps >= 2^2 adjacent columns (128 B runs) inside a 2^tile_log tile
            uint32_t max_len = tile_log - 2;
            uint32_t k = (log_n + max_len - 1) / max_len;
            if (k > NTT_MAX_PASSES) return false;
            plan->npasses = k;
            uint32_t base = log_n / k, extra = log_n % k;
            for (uint32_t i = 0; i < k; i++) plan->log_len[i] = base + (i < extra ? 1 : 0);
        }
    }
    for (uint32_t i = 0; i < plan->npasses; i++) {
        if (plan->log_len[i] > NTT_MAX_TILE_LOG || plan->log_len[i] == 0) {
            if (!(plan->npasses == 1 && log_n == 0)) return false;
        }
    }
    // tile widths
    uint32_t before = 0;
    for (uint32_t i = 0; i < plan->npasses; i++) {
        uint32_t l = plan->log_len[i];
        uint32_t room = (l >= tile_log) ? 0 : tile_log - l;
        if (room > 4) room = 4;                        // up to 16 columns (512 B runs): a 2^6 pass fills a 1024-element tile
        uint32_t avail;
        if (plan->npasses == 1) avail = 0;
        else if (i + 1 < plan->npasses) avail = log_n - before - l;     // log2(columns) = log_stride
        else avail = plan->log_len[0];                                  // last pass: adjacent k0 rows
        plan->log_cw[i] = room < avail ? room : avail;
        before += l;
    }
    return true;
}

