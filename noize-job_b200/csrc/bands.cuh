// bands.cuh — row bands of one grid over the GPUs this process drives (bands.cu); shared with the banded host layer (abi.cu).
#pragma once
#include <functional>
#include <memory>
#include <vector>
#include "nz_common.cuh"

namespace nz {

struct Comm;   // NCCL communicator of one band per process (bands.cu)

struct Band {
    int device = 0, rank = 0;
    int z0 = 0, z1 = 0, own = 0;       // owned rows [z0, z1) of the grid
    int above = 0, below = 0;          // ghost rows that exist above / below the own rows (0 at the grid border)
    float* buf[2] = {nullptr, nullptr};  // ping-pong pair, (above + own + below) x width each
    int cur = 0;                       // which of the two holds the current field
    cudaStream_t s = nullptr;
    bool own_stream = false;
    cudaEvent_t done = nullptr;        // after the band's last stage: neighbours wait for it before they pull ghost rows
    cudaEvent_t pulled = nullptr;      // after the band's last pull: neighbours wait for it before they overwrite a buffer
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // stage timing of the chain
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    int vz0 = 0, vz1 = 0;              // owned vertex rows of the mesh slice
    void* vtx = nullptr;
    uint32_t* idx = nullptr;
    size_t vtx_bytes = 0, idx_bytes = 0;
    long long halo_bytes = 0;          // ghost-row bytes received so far
};

struct BandSet {
    int width = 0, rows = 0, world = 1;
    int cap = 0;                       // ghost-row capacity above and below every band
    std::vector<Band> b;               // the bands this process drives, ascending rank
    Comm* comm = nullptr;              // set: one band per process, neighbours are NCCL ranks
    ~BandSet();
};

using BandStageFn = std::function<int32_t(Band&, float* cur, float* other, int rows, int row_first, float** result)>;

void band_rows(int rows, int world, int rank, int* z0, int* z1);
int32_t bandset_create(std::unique_ptr<BandSet>* out, int width, int rows, int world, int first_rank, int n_local,
                       const int* devices, int cap, Comm* comm, cudaStream_t stream);
int32_t bandset_exchange(BandSet& bs, int above, int below);
int32_t bandset_stage(BandSet& bs, int above, int below, const BandStageFn& fn);
int32_t bandset_sync(BandSet& bs);
int32_t band_scratch(Band& bd, size_t bytes, void** out);
void band_vertex_rows(const Band& bd, int world, int in_res, int R, int* vz0, int* vz1);

int32_t bandset_fractal(BandSet& bs, int noise_type, float hurst, float start_amp, float stepdown, float detune, int octaves,
                        int xpos, int zpos, int noise_size, int extra_above, int extra_below);
int32_t bandset_separable(BandSet& bs, int ksize, const float* kx, const float* kz, float factor, int iterations,
                          int extra_above, int extra_below, bool exchange);
int32_t bandset_sobel2d(BandSet& bs, int iterations, int extra_above, int extra_below, bool exchange);
int32_t bandset_min_erosion(BandSet& bs, int iterations, int extra_above, int extra_below, bool exchange);
bool bandset_flowmap_fits(const BandSet& bs, int iterations);
int32_t bandset_flowmap(BandSet& bs, int iterations, float norm_min, float norm_max, int extra_above, int extra_below, bool exchange);
int32_t bandset_mesh(BandSet& bs, int mesh_type, int R, float tile_height, float tile_size);

}  // namespace nz
