/* Development probe: is the NVIDIA OpenGL driver that sits on the GPU box (libEGL_nvidia.so.0, no GLVND dispatcher, no headers)
 * usable for a headless context?  Bootstraps the vendor library through its GLVND entry point __egl_Main with a minimal set of
 * dispatcher callbacks, then tries EGL_EXT_platform_device -> eglInitialize -> an OpenGL 3.3 core context on a pbuffer.
 * build: gcc -O1 -o tools/glprobe/eglprobe tools/glprobe/eglprobe.c -ldl        run on the box: tools/glprobe/eglprobe */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void* EGLDisplay; typedef void* EGLContext; typedef void* EGLSurface; typedef void* EGLConfig; typedef void* EGLDeviceEXT;
typedef unsigned int EGLenum, EGLBoolean; typedef int EGLint; typedef intptr_t EGLAttrib;
typedef void (*fnptr)(void);
#define EGL_PLATFORM_DEVICE_EXT 0x313F
#define EGL_OPENGL_API 0x30A2
#define EGL_NONE 0x3038

typedef struct VendorInfo { int dummy; } VendorInfo;
typedef struct {
    void (*threadInit)(void);
    EGLenum (*getCurrentApi)(void);
    VendorInfo* (*getCurrentVendor)(void);
    EGLContext (*getCurrentContext)(void);
    EGLDisplay (*getCurrentDisplay)(void);
    EGLSurface (*getCurrentSurface)(EGLint readDraw);
    fnptr (*fetchDispatchEntry)(VendorInfo* vendor, int index);
    void (*setEGLError)(EGLint errorCode);
    EGLBoolean (*setLastVendor)(VendorInfo* vendor);
    VendorInfo* (*getVendorFromDisplay)(EGLDisplay dpy);
    VendorInfo* (*getVendorFromDevice)(EGLDeviceEXT dev);
    void (*setVendorForDevice)(EGLDeviceEXT dev, VendorInfo* vendor);
    void* pad[8];
} ApiExports;
typedef struct {
    EGLDisplay (*getPlatformDisplay)(EGLenum platform, void* nativeDisplay, const EGLAttrib* attrib_list);
    EGLBoolean (*getSupportsAPI)(EGLenum api);
    const char* (*getVendorString)(int name);
    void* (*getProcAddress)(const char* procName);
    void* (*getDispatchAddress)(const char* procName);
    void (*setDispatchIndex)(const char* procName, int index);
    void* more[16];
} ApiImports;

static VendorInfo g_vendor;
static EGLContext g_ctx; static EGLDisplay g_dpy; static EGLSurface g_surf;
static void threadInit(void) {}
static EGLenum getCurrentApi(void) { return EGL_OPENGL_API; }
static VendorInfo* getCurrentVendor(void) { return &g_vendor; }
static EGLContext getCurrentContext(void) { return g_ctx; }
static EGLDisplay getCurrentDisplay(void) { return g_dpy; }
static EGLSurface getCurrentSurface(EGLint rd) { (void)rd; return g_surf; }
static fnptr fetchDispatchEntry(VendorInfo* v, int i) { (void)v; (void)i; return NULL; }
static void setEGLError(EGLint e) { if (e != 0x3000) fprintf(stderr, "[vendor] setEGLError 0x%x\n", e); }
static EGLBoolean setLastVendor(VendorInfo* v) { (void)v; return 1; }
static VendorInfo* getVendorFromDisplay(EGLDisplay d) { (void)d; return &g_vendor; }
static VendorInfo* getVendorFromDevice(EGLDeviceEXT d) { (void)d; return &g_vendor; }
static void setVendorForDevice(EGLDeviceEXT d, VendorInfo* v) { (void)d; (void)v; }

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "/usr/local/nvidia/lib/libEGL_nvidia.so.0";
    void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { printf("dlopen %s: %s\n", path, dlerror()); return 2; }
    typedef EGLBoolean (*MainFn)(uint32_t, const ApiExports*, VendorInfo*, ApiImports*);
    MainFn eglMain = (MainFn)dlsym(h, "__egl_Main");
    if (!eglMain) { printf("no __egl_Main\n"); return 2; }
    ApiExports ex; memset(&ex, 0, sizeof ex);
    ex.threadInit = threadInit; ex.getCurrentApi = getCurrentApi; ex.getCurrentVendor = getCurrentVendor; ex.getCurrentContext = getCurrentContext;
    ex.getCurrentDisplay = getCurrentDisplay; ex.getCurrentSurface = getCurrentSurface; ex.fetchDispatchEntry = fetchDispatchEntry;
    ex.setEGLError = setEGLError; ex.setLastVendor = setLastVendor; ex.getVendorFromDisplay = getVendorFromDisplay;
    ex.getVendorFromDevice = getVendorFromDevice; ex.setVendorForDevice = setVendorForDevice;
    ApiImports im; memset(&im, 0, sizeof im);
    EGLBoolean ok = 0;
    for (uint32_t minor = 0; minor <= 2 && !ok; minor++) { ok = eglMain((0u << 16) | minor, &ex, &g_vendor, &im); printf("__egl_Main(abi 0.%u) -> %u\n", minor, ok); }
    if (!ok) return 3;
    printf("imports: getPlatformDisplay %p getProcAddress %p vendor string %s\n", (void*)im.getPlatformDisplay, (void*)im.getProcAddress,
           im.getVendorString ? im.getVendorString(0) : "(none)");
    fflush(stdout);
    typedef EGLBoolean (*QueryDevicesFn)(EGLint, EGLDeviceEXT*, EGLint*);
    typedef EGLBoolean (*InitializeFn)(EGLDisplay, EGLint*, EGLint*);
    typedef const char* (*QueryStringFn)(EGLDisplay, EGLint);
    typedef EGLint (*GetErrorFn)(void);
    QueryDevicesFn eglQueryDevicesEXT = (QueryDevicesFn)im.getProcAddress("eglQueryDevicesEXT");
    InitializeFn eglInitialize = (InitializeFn)im.getProcAddress("eglInitialize");
    QueryStringFn eglQueryString = (QueryStringFn)im.getProcAddress("eglQueryString");
    GetErrorFn eglGetError = (GetErrorFn)im.getProcAddress("eglGetError");
    printf("  eglQueryDevicesEXT %p eglInitialize %p eglQueryString %p eglGetError %p\n", (void*)eglQueryDevicesEXT, (void*)eglInitialize, (void*)eglQueryString, (void*)eglGetError);
    fflush(stdout);
    EGLDeviceEXT devs[16]; EGLint nd = 0;
    if (eglQueryDevicesEXT) { EGLBoolean q = eglQueryDevicesEXT(16, devs, &nd); printf("eglQueryDevicesEXT -> %u, %d devices, error 0x%x\n", q, nd, eglGetError()); }
    fflush(stdout);
    for (int i = 0; i < nd; i++) {
        EGLDisplay d = im.getPlatformDisplay(EGL_PLATFORM_DEVICE_EXT, devs[i], NULL);
        printf("device %d: display %p\n", i, d); fflush(stdout);
        if (!d) continue;
        EGLint maj = 0, min = 0;
        EGLBoolean r = eglInitialize(d, &maj, &min);
        printf("  eglInitialize -> %u (EGL %d.%d), error 0x%x\n", r, maj, min, eglGetError ? eglGetError() : -1); fflush(stdout);
        if (r) { printf("  vendor %s | version %s | apis %s\n", eglQueryString(d, 0x3053), eglQueryString(d, 0x3054), eglQueryString(d, 0x308D)); g_dpy = d; break; }
    }
    if (!g_dpy) {
        const EGLenum plats[3] = { 0x31DD /* EGL_PLATFORM_SURFACELESS_MESA */, EGL_PLATFORM_DEVICE_EXT, 0x31D7 /* EGL_PLATFORM_GBM_KHR */ };
        for (int k = 0; k < 2 && !g_dpy; k++) {
            EGLDisplay d = im.getPlatformDisplay(plats[k], NULL, NULL);
            printf("platform 0x%x default display: %p (error 0x%x)\n", plats[k], d, eglGetError()); fflush(stdout);
            if (!d) continue;
            EGLint maj = 0, min = 0;
            EGLBoolean r = eglInitialize(d, &maj, &min);
            printf("  eglInitialize -> %u (EGL %d.%d), error 0x%x\n", r, maj, min, eglGetError()); fflush(stdout);
            if (r) { printf("  vendor %s | version %s | apis %s\n", eglQueryString(d, 0x3053), eglQueryString(d, 0x3054), eglQueryString(d, 0x308D)); g_dpy = d; }
        }
    }
    if (!g_dpy) { printf("RESULT: no usable EGL display\n"); return 4; }
    printf("RESULT: EGL display initialised\n");
    return 0;
}
