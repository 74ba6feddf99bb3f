"""The JNI shim (bindings/jni/chaos_jni.c) and the Java classes that call it (bindings/java/).  No JDK exists in the build
image, so what CAN be checked is checked: the C file compiles against a declaration-only stand-in for <jni.h>, every
native method the Java class declares has its JNI function and vice versa, every C-ABI function the shim calls is
declared in include/chaos_ultra.h, and the struct offsets the Java marshalling hard-codes are the header's."""
import re
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
JNI_C = ROOT / "bindings" / "jni" / "chaos_jni.c"
JAVA = ROOT / "bindings" / "java" / "cz" / "cuni" / "mff" / "cgg" / "teichmaa" / "chaosultra" / "b200"


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_jni_shim_compiles_against_the_stand_in_header():
    res = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", str(JNI_C.parent / "compile_check"),
                          "-I", str(ROOT / "include"), str(JNI_C)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_java_natives_and_jni_functions_pair_up():
    java = (JAVA / "ChaosJni.java").read_text()
    natives = set(re.findall(r"static native [\w\[\]]+ (\w+)\(", java))
    c_funcs = set(re.findall(r"JNI_FN\((\w+)\)\(JNIEnv", JNI_C.read_text()))
    assert natives and natives == c_funcs


def test_shim_calls_only_declared_entry_points():
    header = (ROOT / "include" / "chaos_ultra.h").read_text()
    declared = set(re.findall(r"CHAOS_API[^;(]*?\b(chaos_\w+)\s*\(", header))
    called = set(re.findall(r"\b(chaos_[a-z_]+)\(", JNI_C.read_text())) - {"chaos_jobject_"}
    assert called <= declared, called - declared
    for must in ("chaos_provider_create", "chaos_open", "chaos_initialize", "chaos_render_quality", "chaos_render_fast", "chaos_debug",
                 "chaos_set_custom_params", "chaos_supply_defaults", "chaos_close", "chaos_output_rgba", "chaos_driver_display"):
        assert must in called


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_java_marshalling_offsets_are_the_headers(tmp_path):
    src = tmp_path / "o.c"
    src.write_text(r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "chaos_ultra.h"
    int main(void){
      printf("%zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(chaos_params), offsetof(chaos_params, max_iterations), offsetof(chaos_params, segment),
             offsetof(chaos_params, max_super_sampling), offsetof(chaos_params, use_adaptive_super_sampling), offsetof(chaos_params, sample_reuse_cache_dirty),
             offsetof(chaos_params, mouse_focus), offsetof(chaos_params, float_precision));
      printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(chaos_defaults), offsetof(chaos_defaults, has_segment), offsetof(chaos_defaults, center_x),
             offsetof(chaos_defaults, zoom), offsetof(chaos_defaults, max_iterations), offsetof(chaos_defaults, max_super_sampling), offsetof(chaos_defaults, custom_params));
      return 0; }''')
    exe = tmp_path / "o"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    assert got == [64, 4, 8, 40, 44, 50, 52, 60, 296, 4, 8, 24, 32, 36, 40]      # the literals in B200FractalRenderer.java
    java = (JAVA / "B200FractalRenderer.java").read_text()
    assert "PARAMS_BYTES = 64, DEFAULTS_BYTES = 296" in java and "p.getInt(60)" in java and "d.get(40 + i)" in java
