"""The reference-side bindings (bindings/) cannot be compiled here (no JDK, no Scala); what can be checked is checked:
the JNI shim type-checks against include/idprimer.h with a stand-in jni.h, every export of the header has a JNI entry, and
every `@native` the Scala sources declare exists in the shim under the name the JVM will look up."""
import os
import re
import subprocess

import helpers as H

JNI = os.path.join(H.ROOT, "bindings", "jni", "idprimer_jni.c")


def test_jni_shim_type_checks_against_the_header():
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-Wno-unused-parameter", "-I", os.path.join(H.ROOT, "tests", "jni_stub"),
                           "-I", os.path.join(H.ROOT, "include"), JNI])


def test_every_header_export_has_a_jni_entry():
    header = open(os.path.join(H.ROOT, "include", "idprimer.h")).read()
    declared = set(re.findall(r"\b(idp_[a-z0-9_]+)\s*\(", header))
    shim = open(JNI).read()
    missing = sorted(name for name in declared if not re.search(r"\b" + name + r"\s*\(", shim))
    assert not missing, missing


def test_scala_natives_exist_in_the_shim():
    shim = open(JNI).read()
    for src, macro in (("GpuPrimerMatcher.scala", "GPM"), ("AsyncCudaAligner.scala", "ACA")):
        text = open(os.path.join(H.ROOT, "bindings", "scala", src)).read()
        natives = re.findall(r"@native def (\w+)\(", text)
        assert natives, src
        for name in natives:
            assert re.search(macro + r"\(" + name + r"\)", shim), (src, name)
