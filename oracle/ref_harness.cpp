// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the product path.
//
// Thin driver around the UNMODIFIED reference classes (compiled from /root/reference by
// oracle/build_ref.sh into oracle/_ref/).  It replaces the reference's main.cpp so that
//   * ctor / process() / saveResult() can be timed separately (main.cpp:68,111 times all three),
//   * the exit-time crash in ~ImageProcessor (ImageBase.cpp:163, `macroblocks` never initialised
//     by the image ctors) is avoided by leaking the objects and leaving through _exit(),
//   * process() can be repeated for a bounded CPU-baseline measurement.
// It contains no codec logic of its own.
//
// usage:
//   ref_harness enc  <raw> <enc> <W> <H> <rle> <quantfile> [reps]
//   ref_harness dec  <enc> <dec> [reps]
//   ref_harness venc <raw> <enc> <W> <H> <rle> <quantfile> <gop> <merange> [reps]
//   ref_harness vdec <enc> <dec> <motioncomp> [reps]
// Prints one line "@@RESULT {json}" on stderr.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unistd.h>
#include <vector>

#include "main.hpp"
#include "MatrixReader.hpp"
#include "ImageEncoder.hpp"
#include "ImageDecoder.hpp"
#include "VideoEncoder.hpp"
#include "VideoDecoder.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

template <class T> static void run(T *obj, double &t_proc, double &t_save, bool save) {
    double t0 = now_ms();
    obj->process();
    double t1 = now_ms();
    if (save) obj->saveResult();
    double t2 = now_ms();
    t_proc = t1 - t0;
    t_save = t2 - t1;
}

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: see header of ref_harness.cpp\n");
        _exit(64);
    }
    const std::string mode = argv[1];
    int threads = 1;
#ifdef _OPENMP
    threads = omp_get_max_threads();
#endif
    std::vector<double> procs;
    double t_ctor = 0, t_proc = 0, t_save = 0;
    int reps = 1;

    // The reference keeps `const std::string&` members (ImageBase.hpp:42) -> the strings must outlive the objects.
    static std::string a2, a3, a7;
    a2 = argv[2];
    a3 = argv[3];

    if (mode == "enc" || mode == "venc") {
        const bool video = (mode == "venc");
        if (argc < (video ? 10 : 8)) _exit(64);
        static uint16_t W, H, gop = 0, mer = 0;
        static bool rle;
        W = uint16_t(atoi(argv[4]));
        H = uint16_t(atoi(argv[5]));
        rle = atoi(argv[6]) != 0;
        a7 = argv[7];
        if (video) { gop = uint16_t(atoi(argv[8])); mer = uint16_t(atoi(argv[9])); }
        const int ri = video ? 10 : 8;
        if (argc > ri) reps = atoi(argv[ri]);
        static dc::MatrixReader<> m;
        if (!m.read(a7)) _exit(4);
        for (int r = 0; r < reps; r++) {
            double t0 = now_ms();
            if (video) {
                auto *e = new dc::VideoEncoder(a2, a3, W, H, rle, m, gop, mer);
                t_ctor = now_ms() - t0;
                run(e, t_proc, t_save, r == reps - 1);
            } else {
                auto *e = new dc::ImageEncoder(a2, a3, W, H, rle, m);
                t_ctor = now_ms() - t0;
                run(e, t_proc, t_save, r == reps - 1);
            }
            procs.push_back(t_proc);
        }
    } else if (mode == "dec" || mode == "vdec") {
        const bool video = (mode == "vdec");
        static bool mc = true;
        if (video) { if (argc < 5) _exit(64); mc = atoi(argv[4]) != 0; }
        const int ri = video ? 5 : 4;
        if (argc > ri) reps = atoi(argv[ri]);
        for (int r = 0; r < reps; r++) {
            double t0 = now_ms();
            if (video) {
                auto *d = new dc::VideoDecoder(a2, a3, mc);
                t_ctor = now_ms() - t0;
                run(d, t_proc, t_save, r == reps - 1);
            } else {
                auto *d = new dc::ImageDecoder(a2, a3);
                t_ctor = now_ms() - t0;
                run(d, t_proc, t_save, r == reps - 1);
            }
            procs.push_back(t_proc);
        }
    } else {
        _exit(64);
    }

    fflush(stdout);
    fprintf(stderr, "\n@@RESULT {\"mode\": \"%s\", \"block\": %d, \"huffman\": %d, \"threads\": %d, "
                    "\"ctor_ms\": %.3f, \"save_ms\": %.3f, \"process_ms\": [",
            mode.c_str(), int(dc::BlockSize),
#ifdef ENABLE_HUFFMAN
            1,
#else
            0,
#endif
            threads, t_ctor, t_save);
    for (size_t i = 0; i < procs.size(); i++) fprintf(stderr, "%s%.3f", i ? ", " : "", procs[i]);
    fprintf(stderr, "]}\n");
    fflush(stderr);
    _exit(0);   // skip destructors on purpose (see header)
}
