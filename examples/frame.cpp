// examples/frame.cpp — the reference's frame loop (src/main.cpp:69-124) against the C++ mirror,
// headless: renders a few frames of the default scene and writes the last one as a PPM.
//   ./crn_frame [frames] [out.ppm]
#include <cstdio>
#include <cstdlib>

#include "cloud_renderer_b200.hpp"

using namespace crn;

int main(int argc, char **argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 3;
    const char *out = argc > 2 ? argv[2] : "frame.ppm";
    srand(1);                                                       // the reference seeds with time(0)
    try {
        /* Create volume (src/main.cpp:81-82) */
        CloudVolume *volume = new CloudVolume(32, vec2{-5.f, 5.f}, vec3{25.f, 0.f, 0.f}, 4);
        volume->regenerateBillboards(200, vec3{-2.5f, -2.5f, -2.5f}, vec3{2.5f, 2.5f, 2.5f}, 1.f, 2.5f);
        /* Create shaders (src/main.cpp:85-88) */
        VoxelizeShader *voxelizeShader = new VoxelizeShader();
        ConeTraceShader *coneShader = new ConeTraceShader();
        Camera::update();
        for (int f = 0; f < frames; f++) {                          // src/main.cpp:98-124
            Window::runTime = f / 60.f;
            Camera::update();
            Sun::update(volume);
            volume->update();
            voxelizeShader->voxelize(volume);
            coneShader->coneTrace(volume);
        }
        printf("Voxels in scene: %llu\n", (unsigned long long)volume->activeVoxels());   // src/main.cpp:264
        FILE *fp = fopen(out, "wb");
        if (fp) {
            fprintf(fp, "P6\n%d %d\n255\n", Window::width, Window::height);
            for (int y = Window::height - 1; y >= 0; y--)           // row 0 is the bottom row
                for (int x = 0; x < Window::width; x++) fwrite(&coneShader->framebuffer[((size_t)y * Window::width + x) * 4], 1, 3, fp);
            fclose(fp);
            printf("wrote %s\n", out);
        }
        delete coneShader; delete voxelizeShader; delete volume;
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
