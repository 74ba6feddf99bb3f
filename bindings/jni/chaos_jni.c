/*
 * chaos_jni.c -- JNI shim between chaos-ultra's Java host and libchaos_ultra.so (include/chaos_ultra.h).
 *
 * One native method per C-ABI function, class cz.cuni.mff.cgg.teichmaa.chaosultra.b200.ChaosJni (bindings/java/).
 * Handles travel as jlong; the RenderingModel travels as a direct ByteBuffer laid out as struct chaos_params; a failing
 * status becomes the exception the reference throws in the same situation (FractalRenderer.java:14-77,
 * CudaFractalRenderer.java:86,104,160,188, RenderingKernel.java:69,143-146).
 *
 * STATUS: there is no JDK in the build image.  This file is compile-checked against a declaration-only stand-in for
 * <jni.h> (bindings/jni/compile_check/jni.h, tests/test_bindings_cpu.py); it has never been linked against a JVM or run.
 * Build where a JDK exists:
 *   gcc -shared -fPIC -I$JAVA_HOME/include -I$JAVA_HOME/include/linux -I../../include chaos_jni.c -L<dir of libchaos_ultra.so> -lchaos_ultra -o libchaos_jni.so
 */
#include <jni.h>
#include <stdint.h>
#include <string.h>

#include "chaos_ultra.h"

#define JNI_FN(name) Java_cz_cuni_mff_cgg_teichmaa_chaosultra_b200_ChaosJni_##name
#define PROVIDER(h) ((chaos_provider *)(intptr_t)(h))
#define RENDERER(h) ((chaos_renderer *)(intptr_t)(h))

/* chaos_status -> the reference's exception */
static void throw_for(JNIEnv *env, chaos_status st)
{
    static const char *const cls[] = {
        "java/lang/RuntimeException",
        "java/lang/IllegalStateException",
        "java/lang/IllegalArgumentException",
        "cz/cuni/mff/cgg/teichmaa/chaosultra/cudarenderer/CudaInitializationException",
        "cz/cuni/mff/cgg/teichmaa/chaosultra/b200/ChaosCudaException",   /* launch-time error: logged and swallowed unless -Ddebug=true */
        "cz/cuni/mff/cgg/teichmaa/chaosultra/rendering/FractalRendererException",
    };
    jclass c = (*env)->FindClass(env, cls[(int)st >= 1 && (int)st <= 5 ? (int)st : 0]);
    if (c) (*env)->ThrowNew(env, c, chaos_last_error());
}
#define CHECK(call) do { chaos_status st_ = (call); if (st_ != CHAOS_OK) throw_for(env, st_); } while (0)

static chaos_params *params_of(JNIEnv *env, jobject direct_buffer)
{
    chaos_params *p = (chaos_params *)(*env)->GetDirectBufferAddress(env, direct_buffer);
    if (p == NULL || (*env)->GetDirectBufferCapacity(env, direct_buffer) < (jlong)sizeof(chaos_params)) {
        jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
        if (c) (*env)->ThrowNew(env, c, "the model buffer must be a direct ByteBuffer of sizeof(chaos_params) bytes");
        return NULL;
    }
    return p;
}

/* ---- provider: CudaFractalRendererProvider.java:14-91 ---- */
JNIEXPORT jlong JNICALL JNI_FN(providerCreate)(JNIEnv *env, jclass self, jstring kernels_dir, jint device)
{
    (void)self;
    const char *dir = (*env)->GetStringUTFChars(env, kernels_dir, NULL);
    chaos_provider *p = NULL;
    chaos_status st = chaos_provider_create(dir, (int)device, &p);
    (*env)->ReleaseStringUTFChars(env, kernels_dir, dir);
    if (st != CHAOS_OK) { throw_for(env, st); return 0; }
    return (jlong)(intptr_t)p;
}
JNIEXPORT void JNICALL JNI_FN(providerDestroy)(JNIEnv *env, jclass self, jlong provider) { (void)self; CHECK(chaos_provider_destroy(PROVIDER(provider))); }
JNIEXPORT jobjectArray JNICALL JNI_FN(listFractals)(JNIEnv *env, jclass self, jlong provider)
{
    (void)self;
    const char *names[32];
    uint32_t n = 0;
    chaos_status st = chaos_list_fractals(PROVIDER(provider), names, 32, &n);
    if (st != CHAOS_OK) { throw_for(env, st); return NULL; }
    if (n > 32) n = 32;
    jobjectArray out = (*env)->NewObjectArray(env, (jsize)n, (*env)->FindClass(env, "java/lang/String"), NULL);
    for (uint32_t i = 0; out && i < n; ++i) (*env)->SetObjectArrayElement(env, out, (jsize)i, (*env)->NewStringUTF(env, names[i]));
    return out;
}
JNIEXPORT jlong JNICALL JNI_FN(open)(JNIEnv *env, jclass self, jlong provider, jstring name, jboolean force_reload)
{
    (void)self;
    const char *n = (*env)->GetStringUTFChars(env, name, NULL);
    chaos_renderer *r = NULL;
    chaos_status st = chaos_open(PROVIDER(provider), n, force_reload ? 1 : 0, &r);
    (*env)->ReleaseStringUTFChars(env, name, n);
    if (st != CHAOS_OK) { throw_for(env, st); return 0; }
    return (jlong)(intptr_t)r;
}
JNIEXPORT jlong JNICALL JNI_FN(activeRenderer)(JNIEnv *env, jclass self, jlong provider)
{
    (void)env; (void)self;
    return (jlong)(intptr_t)chaos_active_renderer(PROVIDER(provider));
}

/* ---- renderer: FractalRenderer.java:14-77 ---- */
JNIEXPORT void JNICALL JNI_FN(initialize)(JNIEnv *env, jclass self, jlong r, jint width, jint height, jintArray palette)
{
    (void)self;
    jsize len = (*env)->GetArrayLength(env, palette);
    jint *pal = (*env)->GetIntArrayElements(env, palette, NULL);      /* R in the low byte (ImageHelpers.java:138-158) */
    chaos_status st = chaos_initialize(RENDERER(r), (uint32_t)width, (uint32_t)height, (const uint32_t *)pal, (uint32_t)len, CHAOS_OUTPUT_HOST);
    (*env)->ReleaseIntArrayElements(env, palette, pal, JNI_ABORT);
    if (st != CHAOS_OK) throw_for(env, st);
}
JNIEXPORT void JNICALL JNI_FN(freeResources)(JNIEnv *env, jclass self, jlong r) { (void)self; CHECK(chaos_free_resources(RENDERER(r))); }
JNIEXPORT void JNICALL JNI_FN(close)(JNIEnv *env, jclass self, jlong r) { (void)self; CHECK(chaos_close(RENDERER(r))); }
JNIEXPORT void JNICALL JNI_FN(renderQuality)(JNIEnv *env, jclass self, jlong r, jobject model)
{
    (void)self;
    chaos_params *p = params_of(env, model);
    if (p) CHECK(chaos_render_quality(RENDERER(r), p));
}
JNIEXPORT void JNICALL JNI_FN(renderFast)(JNIEnv *env, jclass self, jlong r, jobject model)
{
    (void)self;
    chaos_params *p = params_of(env, model);
    if (p) CHECK(chaos_render_fast(RENDERER(r), p));
}
JNIEXPORT void JNICALL JNI_FN(debug)(JNIEnv *env, jclass self, jlong r) { (void)self; CHECK(chaos_debug(RENDERER(r))); }
JNIEXPORT void JNICALL JNI_FN(setCustomParams)(JNIEnv *env, jclass self, jlong r, jstring text)
{
    (void)self;
    const char *t = (*env)->GetStringUTFChars(env, text, NULL);
    chaos_status st = chaos_set_custom_params(RENDERER(r), t);
    (*env)->ReleaseStringUTFChars(env, text, t);
    if (st != CHAOS_OK) throw_for(env, st);
}
/* supplyDefaultValues: the values come back in a direct buffer laid out as struct chaos_defaults */
JNIEXPORT void JNICALL JNI_FN(supplyDefaults)(JNIEnv *env, jclass self, jlong r, jobject defaults)
{
    (void)self;
    chaos_defaults *d = (chaos_defaults *)(*env)->GetDirectBufferAddress(env, defaults);
    if (d == NULL || (*env)->GetDirectBufferCapacity(env, defaults) < (jlong)sizeof(chaos_defaults)) { throw_for(env, CHAOS_ERR_ILLEGAL_ARGUMENT); return; }
    d->struct_size = (uint32_t)sizeof(chaos_defaults);
    CHECK(chaos_supply_defaults(RENDERER(r), d));
}
JNIEXPORT jint JNICALL JNI_FN(getState)(JNIEnv *env, jclass self, jlong r) { (void)env; (void)self; return (jint)chaos_get_state(RENDERER(r)); }
JNIEXPORT jint JNICALL JNI_FN(getWidth)(JNIEnv *env, jclass self, jlong r) { (void)env; (void)self; return (jint)chaos_get_width(RENDERER(r)); }
JNIEXPORT jint JNICALL JNI_FN(getHeight)(JNIEnv *env, jclass self, jlong r) { (void)env; (void)self; return (jint)chaos_get_height(RENDERER(r)); }
JNIEXPORT jstring JNICALL JNI_FN(fractalName)(JNIEnv *env, jclass self, jlong r) { (void)self; return (*env)->NewStringUTF(env, chaos_fractal_name(RENDERER(r))); }
/* the composed frame: the library's pinned buffer as a direct ByteBuffer, zero copy (row 0 = top, R in the low byte) */
JNIEXPORT jobject JNICALL JNI_FN(outputRgba)(JNIEnv *env, jclass self, jlong r)
{
    (void)self;
    const uint32_t *frame = chaos_output_rgba(RENDERER(r));
    if (!frame) return NULL;
    return (*env)->NewDirectByteBuffer(env, (void *)frame, (jlong)chaos_get_width(RENDERER(r)) * chaos_get_height(RENDERER(r)) * 4);
}

/* ---- frame driver: RenderingModeFSM + the automatic-quality controller, natively (csrc/chaos_driver.cpp) ---- */
JNIEXPORT jlong JNICALL JNI_FN(driverCreate)(JNIEnv *env, jclass self, jlong r, jobject model)
{
    (void)self;
    chaos_params *p = params_of(env, model);
    chaos_driver *d = NULL;
    if (!p) return 0;
    chaos_status st = chaos_driver_create(RENDERER(r), p, &d);
    if (st != CHAOS_OK) { throw_for(env, st); return 0; }
    return (jlong)(intptr_t)d;
}
JNIEXPORT void JNICALL JNI_FN(driverDestroy)(JNIEnv *env, jclass self, jlong d) { (void)self; CHECK(chaos_driver_destroy((chaos_driver *)(intptr_t)d)); }
JNIEXPORT void JNICALL JNI_FN(driverMouse)(JNIEnv *env, jclass self, jlong d, jint x, jint y) { (void)self; CHECK(chaos_driver_mouse((chaos_driver *)(intptr_t)d, x, y)); }
JNIEXPORT void JNICALL JNI_FN(driverStartZooming)(JNIEnv *env, jclass self, jlong d, jboolean inside, jboolean moving)
{
    (void)self;
    CHECK(chaos_driver_start_zooming((chaos_driver *)(intptr_t)d, inside ? 1 : 0, moving ? 1 : 0));
}
JNIEXPORT void JNICALL JNI_FN(driverStopZooming)(JNIEnv *env, jclass self, jlong d) { (void)self; CHECK(chaos_driver_stop_zooming((chaos_driver *)(intptr_t)d)); }
JNIEXPORT void JNICALL JNI_FN(driverStartProgressiveRendering)(JNIEnv *env, jclass self, jlong d, jboolean reset)
{
    (void)self;
    CHECK(chaos_driver_start_progressive_rendering((chaos_driver *)(intptr_t)d, reset ? 1 : 0));
}
/* GLRenderer.display(): 0 nothing rendered, 1 a fast frame, 2 a quality frame */
JNIEXPORT jint JNICALL JNI_FN(driverDisplay)(JNIEnv *env, jclass self, jlong d)
{
    (void)self;
    int kind = 0;
    CHECK(chaos_driver_display((chaos_driver *)(intptr_t)d, &kind));
    return (jint)kind;
}
