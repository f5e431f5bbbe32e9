"""nvcc -c with ptx_fold.py run on the PTX between cicc and ptxas.

    python fold_build.py <nvcc> <src.cu> <out.o> [nvcc flags ...]

nvcc has no hook between its front end and ptxas, so this replays nvcc's own command sequence (`nvcc --dryrun --keep`)
with one command added after cicc.  ptxas' register / spill report goes to stderr as with plain nvcc.  B2_PTX_FOLD=0
compiles with plain nvcc; so does any failure of the replay (with a note on stderr), so the build never depends on it.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def plain(nvcc, src, out, flags):
    return subprocess.call([nvcc] + flags + ["-c", src, "-o", out])


def main():
    nvcc, src, out, flags = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4:]
    if os.environ.get("B2_PTX_FOLD", "1") == "0":
        return plain(nvcc, src, out, flags)
    kd = tempfile.mkdtemp(prefix="b2fold_")
    try:
        dry = subprocess.run([nvcc] + flags + ["--dryrun", "--keep", "--keep-dir", kd, "-c", src, "-o", out], capture_output=True, text=True)
        cmds = [l[3:] for l in dry.stderr.splitlines() if l.startswith("#$ ")]
        if dry.returncode != 0 or not any("cicc" in c for c in cmds):
            raise RuntimeError("nvcc --dryrun failed: " + dry.stderr[-300:])
        script, folded = ["set -e"], 0
        for c in cmds:
            if re.match(r"^\w+=", c):  # nvcc's environment: the commands below only need these three
                if c.split("=", 1)[0] in ("CICC_PATH", "PATH", "LD_LIBRARY_PATH"):
                    script.append("export " + c.strip())
                continue
            script.append(c)
            m = re.search(r'/cicc".*\s-o\s+"([^"]+\.ptx)"', c)
            if m:
                script.append('"%s" "%s" "%s" "%s" 1>&2' % (sys.executable, os.path.join(HERE, "ptx_fold.py"), m.group(1), m.group(1)))
                folded += 1
        if folded != 1:
            raise RuntimeError("expected one cicc step, found %d" % folded)
        path = os.path.join(kd, "build.sh")
        open(path, "w").write("\n".join(script) + "\n")
        rc = subprocess.call(["bash", path])
        if rc != 0:
            raise RuntimeError("replayed command sequence failed (exit %d)" % rc)
        return 0
    except Exception as e:  # noqa: BLE001 -- any trouble: the plain compile is always valid
        sys.stderr.write("fold_build: %s; compiling %s without the PTX pass\n" % (e, src))
        return plain(nvcc, src, out, flags)
    finally:
        shutil.rmtree(kd, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
