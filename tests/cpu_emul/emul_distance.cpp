// emul_distance.cpp -- the distance / playlist kernels of distance.cu run on the host (cuda_on_cpu/cuda_runtime.h),
// on feature rows read from a file; tests/test_host_abi.py compares every output BIT FOR BIT with the oracle (these
// kernels spell their f32 order with __fmul_rn / __fadd_rn, so host and device compute the same bits).
//
//   g++ -std=c++17 -O1 -ffp-contract=off -DBLISS_HOST_EMUL -I tests/cpu_emul/cuda_on_cpu emul_distance.cpp
//   ./emul_distance rows.f32 n dim n_seeds out_dir        (the first n_seeds rows are the seeds)
//
// TEST INFRASTRUCTURE.  The device path sorts the packed keys with cub::DeviceRadixSort; here std::sort stands in for
// it (the keys carry the candidate index in their low half, so any correct sort gives the one stable order).
#include <cuda_runtime.h>

#include <string>

#include "../../bliss-rs_b200/csrc/distance.cu"

using namespace bliss;

static std::string g_out;
template <class T>
static void dump(const std::string &name, const std::vector<T> &v) {
    FILE *f = fopen((g_out + "/" + name).c_str(), "wb");
    if (!f) { perror(name.c_str()); exit(2); }
    fwrite(v.data(), sizeof(T), v.size(), f);
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage: %s rows.f32 n dim n_seeds out_dir\n", argv[0]); return 2; }
    const unsigned n = (unsigned)atoi(argv[2]), dim = (unsigned)atoi(argv[3]), n_seeds = (unsigned)atoi(argv[4]);
    g_out = argv[5];
    std::vector<float> rows((size_t)n * dim);
    {
        FILE *f = fopen(argv[1], "rb");
        if (!f || fread(rows.data(), 4, rows.size(), f) != rows.size()) { perror(argv[1]); return 2; }
        fclose(f);
    }
    // FeaturesVersion::feature_weights (src/lib.rs:168-173, :209-234) as the diagonal the library uploads
    std::vector<float> w(dim, 1.f);
    if (dim == 23) {
        w[0] = 0.25f;
        for (unsigned i = 10; i < 23; i++) w[i] = 3.f / 13.f;
    }
    std::vector<float> full((size_t)dim * dim, 0.f);  // a non-diagonal matrix for the general Mahalanobis form
    for (unsigned i = 0; i < dim; i++)
        for (unsigned j = 0; j < dim; j++) full[(size_t)i * dim + j] = (i == j) ? 1.f + 0.01f * (float)i : 0.001f * (float)((i * 7 + j * 3) % 5);
    dump("full_matrix", full);

    auto matrix = [&](const char *tag, int mode, const float *wm) {
        std::vector<float> out((size_t)n * n, -7.f);
        dim3 grid;
        grid.x = (n + 255u) / 256u;
        grid.y = (n + 31u) / 32u;
        if (dim == 23)
            emu::launch(grid, 128, [&] { distance_matrix_kernel<23>(rows.data(), n, rows.data(), n, mode, wm, out.data()); });
        else if (dim == 20)
            emu::launch(grid, 128, [&] { distance_matrix_kernel<20>(rows.data(), n, rows.data(), n, mode, wm, out.data()); });
        else {
            grid.y = n;
            emu::launch(grid, 256, [&] {
                distance_matrix_generic_kernel(rows.data(), n, rows.data(), n, (int)dim, mode, wm, out.data());
            });
        }
        dump(std::string("matrix_") + tag, out);
    };
    matrix("weights", 0, w.data());     // the crate's default metric
    matrix("euclidean", 0, nullptr);
    matrix("full", 1, full.data());
    matrix("cosine", 2, nullptr);

    // closest_to_songs (src/playlist.rs:256-270): keys = sum of distances to the seeds, stable order
    {
        const float *seeds = rows.data();
        std::vector<float> keys(n, -7.f);
        emu::launch((n + 255u) / 256u, 256, [&] {
            seed_distance_kernel(seeds, n_seeds, rows.data(), n, (int)dim, 0, w.data(), keys.data());
        });
        std::vector<unsigned long long> packed(n);
        emu::launch((n + 255u) / 256u, 256, [&] { make_sort_keys_kernel(keys.data(), n, packed.data()); });
        std::sort(packed.begin(), packed.end());  // cub::DeviceRadixSort::SortKeys on the device
        std::vector<unsigned> order(n);
        emu::launch((n + 255u) / 256u, 256, [&] { unpack_order_kernel(packed.data(), n, order.data()); });
        dump("closest_keys", keys);
        dump("closest_order", order);
    }
    // song_to_song (src/playlist.rs:272-326): greedy nearest-neighbour chain from the seeds
    {
        std::vector<unsigned char> alive(n, 1);
        std::vector<unsigned> order(n, 0xffffffffu);
        std::vector<float> cur(rows.begin(), rows.begin() + (size_t)n_seeds * dim), next(dim, 0.f);
        unsigned n_cur = n_seeds;
        for (unsigned step = 0; step < n; step++) {
            emu::launch(1, 1024, [&] {
                nearest_alive_kernel(cur.data(), n_cur, rows.data(), n, (int)dim, 0, w.data(), alive.data(), order.data(), step,
                                     next.data());
            });
            cur.assign(next.begin(), next.end());
            n_cur = 1;
        }
        dump("chain_order", order);
    }
    printf("OK\n");
    return 0;
}
