// example_chain.cpp — README example #1 of the reference (simplex fBm -> Gauss5 x17 -> FlowMap -> Value Erosion ->
// mesh) written against the C++ stage mirror.  Prints FNV-1a hashes of the heightmap, vertex and index buffers so a
// test can compare them with the same chain driven from the Python mirror (same library => same bits).
//   ./example_chain [resolution]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "noize_stages.hpp"

static unsigned long long fnv(const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    using namespace noize;
    const int res = argc > 1 ? atoi(argv[1]) : 1024, R = res - 8;
    try {
        std::vector<float> tile((size_t)res * res);
        GeneratorData gd;
        gd.uuid = "c2"; gd.resolution = res; gd.xpos = 0; gd.zpos = 0;
        gd.data = nz_slice_f32{tile.data(), 4, (int32_t)tile.size()};

        NoiseStage noise; noise.noiseType = FractalNoise::Simplex; noise.hurst = 0.4f; noise.octaves = 13; noise.noiseSize = 1700;
        KernelFilterStage gauss; gauss.filter = KernelFilterType::Gauss5_S1; gauss.iterations = 17;
        FlowMapStage flow; flow.iterations = 5; flow.normMin = 0.f; flow.normMax = 0.005f;
        ErosionFilterStage erosion; erosion.iterations = 5;
        erosion.keepResident = true;       // the mesh pipeline below works on the same uuid: the tile stays in HBM for it
        BasePipeline generator({&noise, &gauss, &flow, &erosion});
        int completed = 0;
        generator.Run(&gd, [&](StageIO*) { completed++; });

        Mesh mesh;
        MeshStageData md;
        md.uuid = "c2"; md.data = gd.data; md.resolution = R; md.inputResolution = res; md.marginPix = 4;
        md.tileSize = R * (500.0f / 256.0f); md.tileHeight = 2000.f; md.mesh = &mesh;
        MeshTileStage meshStage; meshStage.meshType = MeshType::OvershootSquareGridHeightMap;
        BasePipeline mesher({&meshStage});
        mesher.Run(&md);
        if (GpuResidency::IsOpen("c2")) { printf("ERROR: the mesh stage must have closed the scope\n"); return 3; }

        printf("%s completed=%d launches=%lld height=%016llx vertices=%016llx indices=%016llx\n", nz_version(), completed,
               (long long)nz_kernel_launch_count(), fnv(tile.data(), tile.size() * 4),
               fnv(mesh.vertices.data(), mesh.vertices.size() * sizeof(nz_mesh_vertex)), fnv(mesh.indices.data(), mesh.indices.size() * 4));
        // error behaviour: a wrong StageIO type throws at schedule time, like PipelineStage.cs:37
        try {
            PipelineWorkItem w; w.data = &md;
            noise.Schedule(w, JobHandle());
            printf("ERROR: expected an exception\n");
            return 2;
        } catch (const std::runtime_error&) {}
    } catch (const std::exception& e) {
        fprintf(stderr, "failed: %s\n", e.what());
        return 1;
    }
    return 0;
}
