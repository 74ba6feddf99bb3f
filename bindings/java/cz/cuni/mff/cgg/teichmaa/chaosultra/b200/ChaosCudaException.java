package cz.cuni.mff.cgg.teichmaa.chaosultra.b200;

/** CHAOS_ERR_CUDA: an error at launch or copy time; the reference logs and swallows these unless -Ddebug=true
 *  (cudarenderer/CudaFractalRenderer.java:262-268). */
public class ChaosCudaException extends RuntimeException {
    public ChaosCudaException(String message) {
        super(message);
    }
}
