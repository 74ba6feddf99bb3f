package cz.cuni.mff.cgg.teichmaa.chaosultra.b200;

import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.FractalRenderer;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.FractalRendererState;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.model.DefaultFractalModel;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.model.GLParams;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.model.PlaneSegment;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.model.RenderingModel;
import cz.cuni.mff.cgg.teichmaa.chaosultra.util.FloatPrecision;
import cz.cuni.mff.cgg.teichmaa.chaosultra.util.ImageHelpers;
import cz.cuni.mff.cgg.teichmaa.chaosultra.util.JavaHelpers;

import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.nio.charset.StandardCharsets;

/**
 * rendering/FractalRenderer.java:14-77 over libchaos_ultra.so.  Replaces cudarenderer/CudaFractalRenderer.java; the frame
 * logic (quality / fast, precision rule, double buffer, dirty flags) lives in the library.
 *
 * NOT COMPILED in the build image (no JDK).
 */
public final class B200FractalRenderer implements FractalRenderer {
    /** struct chaos_params (include/chaos_ultra.h): 64 bytes, little endian */
    private static final int PARAMS_BYTES = 64, DEFAULTS_BYTES = 296;
    private final long handle;
    private final ByteBuffer params = ByteBuffer.allocateDirect(PARAMS_BYTES).order(ByteOrder.LITTLE_ENDIAN);
    private GLParams glParams;

    B200FractalRenderer(long handle) {
        this.handle = handle;
    }

    long handle() {
        return handle;
    }

    @Override
    public void initializeRendering(GLParams p) {
        int[] palette = ImageHelpers.loadColorPaletteOrDefault(System.getProperty("colorPalette", "palette.png"));
        ChaosJni.initialize(handle, p.getOutput().getWidth(), p.getOutput().getHeight(), palette);
        glParams = p;
    }

    @Override
    public void freeRenderingResources() {
        ChaosJni.freeResources(handle);
    }

    private ByteBuffer marshal(RenderingModel m) {
        PlaneSegment s = m.getPlaneSegment();
        params.putInt(0, PARAMS_BYTES).putInt(4, m.getMaxIterations());
        params.putDouble(8, s.getLeftBottom().getX()).putDouble(16, s.getLeftBottom().getY());
        params.putDouble(24, s.getRightTop().getX()).putDouble(32, s.getRightTop().getY());
        params.putFloat(40, m.getMaxSuperSampling());
        params.put(44, b(m.isUseAdaptiveSuperSampling())).put(45, b(m.isVisualiseSampleCount()));
        params.put(46, b(m.isUseFoveatedRendering())).put(47, b(m.isUseSampleReuse()));
        params.put(48, b(m.isZooming())).put(49, b(m.isZoomingIn()));
        params.put(50, b(m.isSampleReuseCacheDirty())).put(51, (byte) 0);
        params.putInt(52, m.getMouseFocus().getX()).putInt(56, m.getMouseFocus().getY());
        return params;
    }

    private void render(RenderingModel m, boolean quality) {
        ByteBuffer p = marshal(m);
        try {
            if (quality) ChaosJni.renderQuality(handle, p);
            else ChaosJni.renderFast(handle, p);
        } catch (ChaosCudaException e) {           // CudaFractalRenderer.java:262-268
            m.logError("Error just after launching a kernel:" + e.getMessage());
            if (JavaHelpers.isDebugMode()) throw e;
            return;
        }
        m.setSampleReuseCacheDirty(p.get(50) != 0);
        m.setFloatingPointPrecision(FloatPrecision.values()[p.getInt(60)]);
        // the composed frame is in pinned host memory: one glTexSubImage2D(GL_RGBA, GL_UNSIGNED_BYTE) into
        // glParams.getOutput(), row 0 = top (GLHelpers.java:17-32), replaces the CUDA-GL interop of the reference
        OutputUpload.toTexture(glParams.getOutput(), ChaosJni.outputRgba(handle));
    }

    @Override
    public void renderFast(RenderingModel model) {
        render(model, false);
    }

    @Override
    public void renderQuality(RenderingModel model) {
        render(model, true);
    }

    @Override
    public void launchDebugKernel() {
        ChaosJni.debug(handle);
    }

    @Override
    public void setFractalCustomParams(String text) {
        ChaosJni.setCustomParams(handle, text);
    }

    @Override
    public void supplyDefaultValues(DefaultFractalModel model) {
        ByteBuffer d = ByteBuffer.allocateDirect(DEFAULTS_BYTES).order(ByteOrder.LITTLE_ENDIAN);
        ChaosJni.supplyDefaults(handle, d);
        byte[] text = new byte[256];
        for (int i = 0; i < 256; i++) text[i] = d.get(40 + i);
        int n = 0;
        while (n < 256 && text[n] != 0) n++;
        model.setFractalCustomParams(new String(text, 0, n, StandardCharsets.UTF_8));
        if (d.get(4) != 0) model.setPlaneSegmentFromCenter(d.getDouble(8), d.getDouble(16), d.getDouble(24));
        if (d.get(5) != 0) model.setMaxIterations(d.getInt(32));
        if (d.get(6) != 0) model.setMaxSuperSampling(d.getFloat(36));
    }

    @Override
    public int getWidth() {
        return ChaosJni.getWidth(handle);
    }

    @Override
    public int getHeight() {
        return ChaosJni.getHeight(handle);
    }

    @Override
    public FractalRendererState getState() {
        return ChaosJni.getState(handle) == 1 ? FractalRendererState.readyToRender : FractalRendererState.notInitialized;
    }

    @Override
    public String getFractalName() {
        return ChaosJni.fractalName(handle);
    }

    @Override
    public void close() {
        ChaosJni.close(handle);
    }

    private static byte b(boolean v) {
        return (byte) (v ? 1 : 0);
    }
}
