// main.cpp — the reference's command line (main.go:416-480) in front of the
// CUDA backend: same flags (-S scene, -N cores [ignored by the GPU backend],
// -o output file, -cpuprofile [accepted, ignored]) plus -gpus and -variant.
#include "../../include/grt_host.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char** argv) {
    std::string outFile = "image.ppm";
    int scene = -1, gpus = 1, variant = GRT_VARIANT_AUTO, width = 0, spp = 0;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto val = [&](const char* name) -> const char* {
            std::string f1 = std::string("-") + name, f2 = std::string("--") + name;
            if (a == f1 || a == f2) { if (i + 1 < argc) return argv[++i]; fprintf(stderr, "flag needs an argument: -%s\n", name); exit(2); }
            if (a.rfind(f1 + "=", 0) == 0) return argv[i] + f1.size() + 1;
            if (a.rfind(f2 + "=", 0) == 0) return argv[i] + f2.size() + 1;
            return nullptr;
        };
        const char* v;
        if ((v = val("S"))) scene = atoi(v);
        else if ((v = val("N"))) (void)atoi(v);              // CPU thread count: meaningless for the GPU backend
        else if ((v = val("o")) || (v = val("outfile"))) outFile = v;   // main.go:419 defines -o; the README says -outfile
        else if ((v = val("cpuprofile"))) (void)v;
        else if ((v = val("gpus"))) gpus = atoi(v);
        else if ((v = val("variant"))) variant = (strcmp(v, "wavefront") == 0) ? GRT_VARIANT_WAVEFRONT : (strcmp(v, "mega") == 0 || strcmp(v, "megakernel") == 0) ? GRT_VARIANT_MEGAKERNEL : GRT_VARIANT_AUTO;
        else if ((v = val("width"))) width = atoi(v);
        else if ((v = val("spp"))) spp = atoi(v);
        else { fprintf(stderr, "flag provided but not defined: %s\nUsage: -S int -N int -o string -cpuprofile string [-gpus int] [-variant mega|wavefront]\n", a.c_str()); return 2; }
    }
    FILE* f = fopen(outFile.c_str(), "wb");
    if (!f) { fprintf(stderr, "Error creating output file\n"); return 1; }   // main.go:436-438
    GrtHostScene* s = grt_host_scene_new();
    GrtCameraConfig cam;
    GrtSceneOptions so;
    memset(&so, 0, sizeof(so));
    so.width = width; so.spp = spp;
    if (grt_host_builtin_scene(s, scene, &so, &cam)) {
        // defaultScene (main.go:412) renders nothing and writes an empty file
        fclose(f);
        grt_host_scene_free(s);
        return 0;
    }
    printf("Beginning render. . .\n");   // camera.go:157
    int h = (int)(cam.Width / (cam.AspectRatio == 0 ? 1.0 : cam.AspectRatio));
    if (h < 1) h = 1;
    long cap = 32 + (long)cam.Width * h * 12;
    std::vector<char> ppm((size_t)cap);
    std::vector<uint8_t> rgb8((size_t)cam.Width * h * 3);
    long len = 0;
    double ms = 0;
    // the reference writes P3 text whatever the file is called; ".png" / ".pnm" ask for the binary containers instead
    auto ends_with = [&](const char* e) { size_t n = strlen(e); return outFile.size() >= n && outFile.compare(outFile.size() - n, n, e) == 0; };
    const bool png = ends_with(".png"), p6 = ends_with(".pnm");
    int rc = grt_host_camera_render_rgb8(s, &cam, 0xC0FFEEull, variant, gpus, nullptr, rgb8.data(), (png || p6) ? nullptr : ppm.data(), cap, &len, &ms);
    if (rc) { fprintf(stderr, "render failed (%d): %s / %s\n", rc, grt_host_last_error(), grt_last_error()); fclose(f); return 1; }
    if (getenv("GRT_VERBOSE")) fprintf(stderr, "[grt_main] %d GPU(s): render + exchange %.2f ms (device time, max over devices)\n", gpus, ms);
    if (png || p6) {
        std::vector<unsigned char> bin(rgb8.size() + rgb8.size() / 1000 + (size_t)h * 8 + 256);
        len = png ? grt_host_write_png(rgb8.data(), cam.Width, h, bin.data(), (long)bin.size()) : grt_host_write_p6(rgb8.data(), cam.Width, h, bin.data(), (long)bin.size());
        if (len < 0) { fprintf(stderr, "image encode failed\n"); fclose(f); return 1; }
        fwrite(bin.data(), 1, (size_t)len, f);
    } else
    fwrite(ppm.data(), 1, (size_t)len, f);   // main.go:479
    fclose(f);
    grt_host_scene_free(s);
    return 0;
}
