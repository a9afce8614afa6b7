// ref_wrap_retina_host.cpp -- TEST INFRASTRUCTURE.  CPU-only wrapper around retinaface/common.hpp nms().
#include <cstring>
#include <vector>

#include "common.hpp"

extern "C" {
__attribute__((visibility("default"))) int ref_retina_nms(float* output_host, float nms_thresh, float* res_out) {
    std::vector<decodeplugin::Detection> res;
    nms(res, output_host, nms_thresh);
    for (size_t i = 0; i < res.size(); ++i) memcpy(res_out + i * 15, &res[i], sizeof(decodeplugin::Detection));
    return (int)res.size();
}
}
