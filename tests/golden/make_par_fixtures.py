"""Copies three of the reference's par files (input-format fixtures, not source code) from
/root/reference/etc/SeekDeepParameterFiles into tests/golden/ so the parser tests run where the
reference tree is absent (the GPU box).  Run once in the authoring container."""
import os
import shutil

SRC = "/root/reference/etc/SeekDeepParameterFiles"
DST = os.path.dirname(os.path.abspath(__file__))
for name in ("illumina_lkmer1", "454_it_lkmer2_largeHpr", "otu_99"):
    shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name + ".par"))
    print("wrote", name + ".par")
