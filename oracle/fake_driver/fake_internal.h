/* fake_internal.h — shared between fake_gpu.c (bookkeeping, NVML, symbol table) and fake_exec.c (functional layer).
 * TEST INFRASTRUCTURE ONLY (oracle/). */
#ifndef FAKE_INTERNAL_H
#define FAKE_INTERNAL_H
#include <stddef.h>
#include <stdint.h>

typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef void *CUcontext;
typedef void *CUstream;
typedef void *CUfunction;
typedef void *CUmodule;
typedef void *CUevent;

#define CUDA_SUCCESS 0
#define CUDA_ERROR_INVALID_VALUE 1
#define CUDA_ERROR_OUT_OF_MEMORY 2
#define CUDA_ERROR_NOT_INITIALIZED 3
#define CUDA_ERROR_INVALID_DEVICE 101
#define CUDA_ERROR_INVALID_CONTEXT 201
#define CUDA_ERROR_NOT_FOUND 500
#define CUDA_ERROR_NOT_SUPPORTED 801

/* fake_gpu.c */
int fake_exec_on(void);                                     /* FAKE_GPU_EXEC=1 */
CUresult fake_track(uint64_t base, uint64_t size, int dev); /* record a live allocation + charge the device */
CUresult fake_untrack(uint64_t base, uint64_t *size);
int fake_charge(int dev, int64_t delta);                    /* nonzero = device total exceeded (nothing charged) */

/* fake_exec.c */
CUresult fx_alloc(CUdeviceptr *p, size_t n, int dev);
CUresult fx_free(CUdeviceptr p);
CUresult fx_address_reserve(CUdeviceptr *p, size_t size, size_t align, CUdeviceptr addr, unsigned long long flags);
CUresult fx_address_free(CUdeviceptr p, size_t n);
CUresult fx_mem_create(unsigned long long *h, size_t n, const void *prop, unsigned long long flags);
CUresult fx_mem_release(unsigned long long h);
CUresult fx_mem_map(CUdeviceptr va, size_t n, size_t off, unsigned long long h, unsigned long long flags);
CUresult fx_mem_set_access(CUdeviceptr va, size_t n, const void *desc, size_t cnt);
CUresult fx_mem_unmap(CUdeviceptr va, size_t n);
int fx_is_mapped(CUdeviceptr p);
int fake_is_tracked(uint64_t p);                            /* inside a live plain allocation */
CUresult fx_event_create(CUevent *e);
CUresult fx_event_record(CUevent e);
CUresult fx_event_elapsed(float *ms, CUevent a, CUevent b);
CUresult fx_event_destroy(CUevent e);
CUresult fx_get_function(CUfunction *f, const char *name);
CUresult fx_param_info(CUfunction f, size_t idx, size_t *off, size_t *size);
CUresult fx_launch(CUfunction f, void **params);
#endif
