// scan.cuh -- the scan / session handle shared by pam_scan.cu (K1) and session.cu (the device-resident chain
// scan -> seed flags -> distinct-guide table -> index -> kNN).
#pragma once
#include "common.cuh"

namespace gm {

struct Scan {
    // rows of the PAM scan, resident in HBM
    uint64_t *guides = nullptr;     // guide2bit per row
    uint32_t *start = nullptr;      // 0-based start of the target window (record-relative for sessions)
    uint16_t *pamcode = nullptr;    // exact PAM, 2 bits per base
    int32_t *rec = nullptr;         // sessions: record of the row
    uint8_t *strand = nullptr;      // sessions: 1 = forward
    int64_t n_fwd = 0, n_rev = 0;
    // sessions keep the genome and the scan parameters
    bool session = false;
    uint8_t *seq = nullptr;         // ASCII genome as passed in (records joined by one invalid byte)
    int64_t n_seq = 0;
    int64_t *rec_start = nullptr;   // device copy of the n_rec + 1 record offsets
    int n_rec = 0;
    int P = 0, L = 0, five_prime = 0;
    int32_t *first32 = nullptr;     // after gm_session_index: row -> first row with the same guide
    // after gm_session_neighbors: the kept query rows (compact), waiting for gm_session_fetch_neighbors
    uint64_t *nb_codes = nullptr;
    int32_t *nb_idx = nullptr;
    uint8_t *nb_dist = nullptr;
    int64_t nb_rows = 0;
    int nb_k = 0;
};

// dedup.cu
int dedup_dev(const uint64_t *d_keys, int64_t n, int L, int lsr, int five_prime, uint8_t *d_is_dup, int64_t *d_first_row,
              int32_t *d_first32, cudaStream_t st);
// restriction.cu
int restriction_dev(const uint64_t *d_guides, int64_t n, int L, const uint8_t *motif_sets, const int32_t *motif_len, int n_motifs,
                    uint8_t *d_has_site, cudaStream_t st);

}  // namespace gm
