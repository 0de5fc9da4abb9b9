/*
 * algorithmModule.h -- pthread entry point of the Algorithm stage
 * (drop-in for reference Youth.Source/AlgorithmModule/algorithmModule.h:6; the
 * intended launch site is reference main.c:279-281).
 */
#ifndef ALGORITHM_MODULE_H
#define ALGORITHM_MODULE_H

#include "SLAM.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Thread start routine.  `id` is passed through untouched by the reference; here a
 * non-NULL `id` is read as `const char*` path of a .bin recording to replay
 * headless, NULL means "serve processSlamFrame() callers
 * until stopSlamModule()".  Returns NULL. */
void* algorithmModule(void* id);

#ifdef __cplusplus
}
#endif

#endif /* ALGORITHM_MODULE_H */
