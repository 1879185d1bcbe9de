"""The shims other than ORBextractor.cc cannot be compiled in this image (Eigen, Sophus' dependencies, g2o's, boost,
Pangolin headers are absent), so this is the next best thing to a syntax check: every member / method the shims reach
through `->`, `.` or `Class::` on a reference object must exist, by name, in the reference's own headers -- it catches
a misspelt or renamed member (`NLeft` vs `Nleft`, `mvuRight`, `GetRelativePoseTrl` ...) before a maintainer's compiler
does.  Skipped where the reference tree is absent (GPU box)."""
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

# names that belong to the standard library, Eigen (un-vendored), to this repo's C ABI structs, or to locals of the shims
STD = set("""size push_back data insert end begin empty clear reserve resize first second count find at get reset assign
emplace_back back front erase c_str what lock unlock swap fill max min x y z w cast cols rows ptr pt octave angle response
step total create release isContinuous clone copyTo type str append substr length join emplace cbegin cend
getMat rowRange colRange""".split())
# accessors INTEGRATION.md asks the maintainer to add to the reference headers (each must be named there)
ADDED = {"flatten", "GetMinDistanceRaw", "GetMaxDistanceRaw"}


def _reference_words():
    words = set()
    pats = [os.path.join(REF, "include", "*.h"), os.path.join(REF, "include", "CameraModels", "*.h"),
            os.path.join(REF, "Thirdparty", "Sophus", "sophus", "*.hpp"),
            os.path.join(REF, "Thirdparty", "DBoW2", "DBoW2", "*.h"), os.path.join(REF, "Thirdparty", "g2o", "g2o", "*", "*.h")]
    for pat in pats:
        for f in glob.glob(pat):
            words.update(re.findall(r"[A-Za-z_]\w*", open(f, errors="ignore").read()))
    return words


def _own_words():
    words = set()
    for f in glob.glob(os.path.join(ROOT, "include", "*.h")) + glob.glob(os.path.join(ROOT, "orb_slam3_b200", "shim", "*.h")):
        words.update(re.findall(r"[A-Za-z_]\w*", open(f).read()))
    return words


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "include", "Frame.h")), reason="reference tree not present")
def test_every_member_the_shims_touch_exists_in_the_reference_headers():
    ref, own = _reference_words(), _own_words()
    missing = {}
    shims = sorted(glob.glob(os.path.join(ROOT, "orb_slam3_b200", "shim", "*.cc")) +
                   glob.glob(os.path.join(ROOT, "orb_slam3_b200", "shim", "*.h")))
    assert len(shims) >= 9
    for path in shims:
        src = open(path).read()
        src = re.sub(r"//[^\n]*", "", src)                      # comments cite reference lines in prose
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r'"(\\.|[^"\\])*"', '""', src)               # string literals
        # a name that also occurs on its own (a declaration, a local variable, a field of a struct the shim defines) is
        # the shim's; a member of a reference object only ever follows `->`, `.` or `::`
        local = set(re.findall(r"(?<![>.:\w])([A-Za-z_]\w*)\b(?!\s*\()", src))
        names = set(re.findall(r"(?:->|\.)\s*([A-Za-z_]\w*)", src))
        names.update(re.findall(r"\b(?:Frame|KeyFrame|MapPoint|Map|Optimizer|ORBmatcher|ORBextractor|Tracking|GeometricCamera|"
                                r"Converter|Verbose|Sophus|Eigen|g2o|DBoW2)::([A-Za-z_]\w*)", src))
        for name in sorted(names):
            if name in STD or name in own or name in local:
                continue
            if name in ADDED:
                assert name + "(" in open(os.path.join(ROOT, "INTEGRATION.md")).read(), name
                continue
            if name not in ref:
                missing.setdefault(os.path.basename(path), []).append(name)
    assert not missing, missing


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "include", "Frame.h")), reason="reference tree not present")
def test_replaced_signatures_are_the_reference_ones():
    """The method definitions in the shims carry exactly the parameter lists the reference headers declare."""
    def norm(sig):
        sig = re.sub(r"=\s*[^,)]+", "", sig)                      # default arguments appear in the header only
        sig = re.sub(r"\bstd::", "", sig)
        sig = re.sub(r"\b(\w+)\s*(?=[,)])", "", sig)              # parameter names
        return re.sub(r"\s+", "", sig).replace("const", "")
    decls = {
        "Optimizer.h": ["LocalBundleAdjustment", "PoseOptimization", "LocalInertialBA"],
        "ORBmatcher.h": ["SearchForTriangulation", "DescriptorDistance"],
        "Frame.h": ["ComputeStereoMatches", "ComputeBoW"],
    }
    shim_src = "\n".join(open(f).read() for f in glob.glob(os.path.join(ROOT, "orb_slam3_b200", "shim", "*.cc")))
    for header, methods in decls.items():
        h = open(os.path.join(REF, "include", header)).read()
        for meth in methods:
            hm = re.search(r"\b%s\s*\(([^;{]*)\)\s*(?:const)?\s*;" % meth, h)
            sm = re.search(r"\b\w+::%s\s*\(([^;{]*)\)\s*(?:const)?\s*\{" % meth, shim_src)
            if not sm:      # a sketch that only holds fragments of the method does not define it
                continue
            assert hm, (header, meth)
            assert norm("(" + hm.group(1) + ")") == norm("(" + sm.group(1) + ")"), (meth, hm.group(1), sm.group(1))
