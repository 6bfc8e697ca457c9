// internal.h -- what the other host-side translation units of libdwtb200.so (strips.cu) need from dwtb200.cu.
#pragma once
#include <cstdint>
#include <mutex>
#include <cuda_runtime.h>

struct dwtb200_image;

namespace dwtb200 {

struct ImageView {
    void *plane[2];
    int cur;
    int64_t pitch, frame;   // elements
    size_t es;
    int ox, oy, frames, kind;
    cudaStream_t st;        // the image's own stream: every call on the image is ordered on it
};
ImageView image_view(const dwtb200_image *im);
int set_error(int code, const char *fmt, ...);   // records the message of dwtb200_last_error(), returns code
std::recursive_mutex &api_mutex();               // the process-wide lock every extern "C" entry point takes

}  // namespace dwtb200
