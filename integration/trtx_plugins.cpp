// The whole plugin library of a tensorrtx model directory (replaces plugin/yololayer.cu, decode.cu, rcnn/*Plugin.h):
// static registrars for "YoloLayer_TRT", "Decode_TRT", "RpnDecode", "RpnNms", "PredictorDecode", "BatchedNms" (version "1").
// Link with libtrtx_hot.so, nvinfer and cudart.  See INTEGRATION.md.
#define TRTX_REGISTER_PLUGINS
// #define TRTX_IMPLICIT_BATCH_PLUGINS   // for builders that still use setMaxBatchSize(): IPluginV2IOExt objects
#include "trtx_plugins.h"
