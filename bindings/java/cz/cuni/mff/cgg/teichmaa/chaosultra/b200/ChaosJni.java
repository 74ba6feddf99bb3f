package cz.cuni.mff.cgg.teichmaa.chaosultra.b200;

import java.nio.ByteBuffer;

/**
 * Native methods of bindings/jni/chaos_jni.c, one per function of include/chaos_ultra.h that the Java host needs.
 * Handles are pointers carried as long; the model travels as a direct ByteBuffer laid out as struct chaos_params.
 *
 * NOT COMPILED in the build image (no JDK): see INTEGRATION.md.  tests/test_bindings_cpu.py checks that every native
 * method declared here has its JNI function in chaos_jni.c and vice versa.
 */
final class ChaosJni {
    static {
        System.loadLibrary("chaos_jni");   // which links libchaos_ultra.so
    }

    private ChaosJni() {
    }

    static native long providerCreate(String kernelsDir, int device);
    static native void providerDestroy(long provider);
    static native String[] listFractals(long provider);
    static native long open(long provider, String fractalName, boolean forceReload);
    static native long activeRenderer(long provider);

    static native void initialize(long renderer, int width, int height, int[] paletteRgba);
    static native void freeResources(long renderer);
    static native void close(long renderer);
    static native void renderQuality(long renderer, ByteBuffer model);
    static native void renderFast(long renderer, ByteBuffer model);
    static native void debug(long renderer);
    static native void setCustomParams(long renderer, String text);
    static native void supplyDefaults(long renderer, ByteBuffer defaults);
    static native int getState(long renderer);
    static native int getWidth(long renderer);
    static native int getHeight(long renderer);
    static native String fractalName(long renderer);
    static native ByteBuffer outputRgba(long renderer);

    static native long driverCreate(long renderer, ByteBuffer model);
    static native void driverDestroy(long driver);
    static native void driverMouse(long driver, int x, int y);
    static native void driverStartZooming(long driver, boolean inside, boolean movingToo);
    static native void driverStopZooming(long driver);
    static native void driverStartProgressiveRendering(long driver, boolean resetFirst);
    static native int driverDisplay(long driver);
}
