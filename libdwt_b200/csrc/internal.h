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
// row strips: level 0 of a forward transform reads rows [0, up_end) / [dn_begin, ...) of the image from other planes (one pointer per
// plane of the image: the transform alternates between them); nullptr arrays switch it off.  Drops the image's cached graphs.
void image_set_row_sources(dwtb200_image *im, const void *const up[2], const void *const dn[2], int up_end, int dn_begin, int64_t up_row0,
                           int64_t dn_row0);
bool image_level0_is_ring(dwtb200_image *im, int J);
int set_error(int code, const char *fmt, ...);   // records the message of dwtb200_last_error(), returns code
std::recursive_mutex &api_mutex();               // the process-wide lock every extern "C" entry point takes

}  // namespace dwtb200
