package cz.cuni.mff.cgg.teichmaa.chaosultra.b200;

import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.FractalRenderer;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.FractalRendererProvider;

import java.io.File;
import java.util.Arrays;
import java.util.LinkedHashSet;
import java.util.Set;

/**
 * rendering/FractalRendererProvider.java:11-27 over libchaos_ultra.so; replaces
 * cudarenderer/CudaFractalRendererProvider.java:14-91.  rendering/GLRenderer.java:81 becomes
 * {@code new B200FractalRendererProvider()} -- the only edit inside the existing sources.
 *
 * NOT COMPILED in the build image (no JDK).
 */
public final class B200FractalRendererProvider implements FractalRendererProvider {
    private final long provider;
    private B200FractalRenderer active;

    public B200FractalRendererProvider() {
        // cudarenderer/FractalRenderingModule.java:38-55: <user.dir>/<-DcudaKernelsDir, default cudaKernels>
        String dir = System.getProperty("user.dir") + File.separator + System.getProperty("cudaKernelsDir", "cudaKernels");
        provider = ChaosJni.providerCreate(dir, Integer.getInteger("cudaDevice", 0));
    }

    @Override
    public Set<String> getAvailableFractals() {
        return new LinkedHashSet<>(Arrays.asList(ChaosJni.listFractals(provider)));
    }

    @Override
    public FractalRenderer getDefaultRenderer() {
        return getRenderer("mandelbrot", false);
    }

    @Override
    public FractalRenderer getRenderer(String fractalName, boolean forceReload) {
        long before = active == null ? 0 : active.handle();
        long h;
        try {
            h = ChaosJni.open(provider, fractalName, forceReload);
        } finally {
            if (ChaosJni.activeRenderer(provider) != before) active = null;   // the library closed the previous renderer (:52)
        }
        if (active == null || active.handle() != h) active = new B200FractalRenderer(h);
        return active;
    }

    public void close() {
        ChaosJni.providerDestroy(provider);
    }
}
