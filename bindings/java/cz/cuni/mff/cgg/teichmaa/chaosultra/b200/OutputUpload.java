package cz.cuni.mff.cgg.teichmaa.chaosultra.b200;

import com.jogamp.opengl.GL2;
import com.jogamp.opengl.GLContext;
import cz.cuni.mff.cgg.teichmaa.chaosultra.rendering.model.GLTexture;

import java.nio.ByteBuffer;

/** The composed frame (pinned host memory, RGBA8, row 0 = top) into the output texture GLRenderer draws
 *  (rendering/GLHelpers.java:17-32 creates it as GL_RGBA / GL_UNSIGNED_BYTE).  NOT COMPILED in the build image. */
final class OutputUpload {
    private OutputUpload() {
    }

    static void toTexture(GLTexture texture, ByteBuffer rgba) {
        GL2 gl = GLContext.getCurrentGL().getGL2();
        gl.glBindTexture(GL2.GL_TEXTURE_2D, texture.getHandle().getValue());
        gl.glTexSubImage2D(GL2.GL_TEXTURE_2D, 0, 0, 0, texture.getWidth(), texture.getHeight(), GL2.GL_RGBA, GL2.GL_UNSIGNED_BYTE, rgba);
        gl.glBindTexture(GL2.GL_TEXTURE_2D, 0);
    }
}
