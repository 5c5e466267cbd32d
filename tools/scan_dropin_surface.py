"""Which members of the reference's StateValidityChecker family do its planners and run/ drivers actually call?

For every flavour (PCS, GD, HB, RX, SG = RSS checker) this scans the reference's planners/*_<F>*.cpp, planners/*_<F>.h and
run/plan_<F>*.cpp for calls (unqualified, `this->` or `StateValidityChecker::` / base-class qualified) to any member the
reference's checker class or its bases (two_robots / kdl / collisionDetection) declare, and records the name with the
argument counts used.  Output: tests/golden/dropin_surface.json (committed), which tests/test_abi_cpu.py holds the drop-in
headers of abb_dualarm_planning_b200/host/ to -- so "the planners compile unchanged" is checked mechanically.

    python tools/scan_dropin_surface.py            (needs /root/reference)
"""
import glob
import json
import os
import re
import sys

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

FLAVOURS = {
    # flavour: (checker header, base-class headers, planner suffixes)
    "PCS": ("validity_checkers/StateValidityCheckerPCS.h", ["proj_classes/apc_class.h", "validity_checkers/collisionDetection.h"], ["PCS", "PCS_s"]),
    "GD": ("validity_checkers/StateValidityCheckerGD.h", ["proj_classes/kdl_class.h", "validity_checkers/collisionDetection.h"], ["GD", "GD_s"]),
    "HB": ("validity_checkers/StateValidityCheckerHB.h", ["proj_classes/apc_class.h", "proj_classes/kdl_class.h", "validity_checkers/collisionDetection.h"], ["HB"]),
    "RX": ("validity_checkers/StateValidityCheckerRX.h", ["proj_classes/kdl_class.h", "validity_checkers/collisionDetection.h"], ["RX"]),
    "SG": ("validity_checkers/StateValidityCheckerRSS.h", ["proj_classes/apc_class.h", "validity_checkers/collisionDetection.h"], ["SG"]),
}
KEYWORDS = {"if", "for", "while", "switch", "return", "sizeof", "catch", "defined", "assert"}


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    s = re.sub(r"//[^\n]*", " ", s)
    s = re.sub(r'"(?:\\.|[^"\\])*"', '""', s)
    return s


def split_args(argtext, angle=False):
    """top-level comma split of an argument / parameter list (angle: '<' '>' nest, for template types in declarations)"""
    out, depth, cur = [], 0, ""
    opens, closes = ("([{<", ")]}>") if angle else ("([{", ")]}")
    for ch in argtext:
        if ch in opens:
            depth += 1
        elif ch in closes:
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def matching_paren(s, i):
    """s[i] == '(' -> index of the matching ')'"""
    depth = 0
    for j in range(i, len(s)):
        if s[j] == "(":
            depth += 1
        elif s[j] == ")":
            depth -= 1
            if depth == 0:
                return j
    return -1


def declared_members(header_text):
    """{name: [(min_args, max_args), ...]} of the function-like declarations in a header (any class scope)"""
    s = strip_comments(header_text)
    out = {}
    for m in re.finditer(r"([A-Za-z_]\w*)\s*\(", s):
        name = m.group(1)
        if name in KEYWORDS:
            continue
        # a declaration is preceded by a type-ish token (identifier, '>', '&', '*') and followed, after the ')', by ; { const or :
        pre = s[:m.start()].rstrip()
        if not pre or not re.search(r"[\w>&*~]$", pre) or re.search(r"(return|else|new|delete|=)$", pre):
            continue
        j = matching_paren(s, m.end() - 1)
        if j < 0:
            continue
        post = s[j + 1:j + 40].lstrip()
        if not re.match(r"(const\b)?\s*(;|\{|:|override|=\s*0)", post):
            continue
        params = [p for p in split_args(s[m.end():j], angle=True) if p.strip() and p.strip() != "void"]
        nmax = len(params)
        nmin = sum(1 for p in params if "=" not in p)
        out.setdefault(name, [])
        if (nmin, nmax) not in out[name]:
            out[name].append((nmin, nmax))
    return out


def calls_in(text, names):
    """{name: sorted arg counts} of the calls to any of `names` that are not made on another object"""
    s = strip_comments(text)
    out = {}
    for m in re.finditer(r"([A-Za-z_]\w*)\s*\(", s):
        name = m.group(1)
        if name not in names:
            continue
        pre = s[:m.start()].rstrip()
        if pre.endswith(".") or (pre.endswith("->") and not pre.endswith("this->")):
            continue                          # a call on some other object (si_->..., pdef_->..., a vector, ...)
        if re.search(r"(\w|>|&|\*)$", pre) and not re.search(r"(return|else|&&|\|\||!)$", pre) and not pre.endswith("::"):
            # "Type name(" = a declaration / definition, not a call (e.g. the planner's own member definitions)
            if re.search(r"[A-Za-z_]\w*\s*[&*]*$", pre) and not re.search(r"(return|else|case)\s*$", pre):
                continue
        if pre.endswith("::"):
            q = re.search(r"([A-Za-z_]\w*)\s*::$", pre).group(1)
            if q not in ("StateValidityChecker", "two_robots", "kdl", "collisionDetection"):
                continue                      # Planner::member definition or another namespace
        j = matching_paren(s, m.end() - 1)
        if j < 0:
            continue
        nargs = len([a for a in split_args(s[m.end():j]) if a.strip()])
        out.setdefault(name, set()).add(nargs)
    return {k: sorted(v) for k, v in out.items()}


def scan_reference():
    res = {}
    for flav, (hdr, bases, sufs) in FLAVOURS.items():
        decl = {}
        for h in [hdr] + bases:
            for k, v in declared_members(open(os.path.join(REF, h), errors="replace").read()).items():
                decl.setdefault(k, []).extend(v)
        # public data members the planners read directly (counters)
        files = []
        for suf in sufs:
            files += glob.glob(os.path.join(REF, "planners", "*_%s.cpp" % suf)) + glob.glob(os.path.join(REF, "planners", "*_%s.h" % suf))
            files += glob.glob(os.path.join(REF, "run", "plan_%s*.cpp" % suf))
        used = {}
        for f in sorted(set(files)):
            own = declared_members(open(f[:-4] + ".h", errors="replace").read()) if os.path.exists(f[:-4] + ".h") else {}
            names = {k for k in decl if k not in own and k not in ("StateValidityChecker", "two_robots", "kdl", "collisionDetection")}
            for k, v in calls_in(open(f, errors="replace").read(), names).items():
                e = used.setdefault(k, {"arities": set(), "files": set()})
                e["arities"].update(v); e["files"].add(os.path.relpath(f, REF))
        res[flav] = {k: {"arities": sorted(v["arities"]), "files": sorted(v["files"])[:4],
                         "reference_signatures": sorted(set(decl[k]))} for k, v in sorted(used.items())}
    return res


OURS = {  # flavour -> (header, extra -D)
    "PCS": "StateValidityCheckerPCS.h", "GD": "StateValidityCheckerGD.h", "HB": "StateValidityCheckerHB.h",
    "RX": "StateValidityCheckerRX.h", "SG": "StateValidityCheckerRSS.h",
}


def our_members(flavour):
    """declarations visible through the drop-in header of a flavour: the header is run through the preprocessor (so the
    CKC_FLAVOUR_* blocks are resolved) and only the text that comes from files of host/ is kept"""
    import subprocess
    host = os.path.join(REPO, "abb_dualarm_planning_b200", "host")
    pre = subprocess.run(["g++", "-std=c++14", "-E", os.path.join(host, OURS[flavour])], capture_output=True, text=True, check=True).stdout
    keep, text = False, []
    for line in pre.split("\n"):
        m = re.match(r'# \d+ "([^"]*)"', line)
        if m:
            keep = os.path.dirname(os.path.abspath(m.group(1))) == host
            continue
        if keep:
            text.append(line)
    return declared_members("\n".join(text))


def missing(surface, flavour):
    ours = our_members(flavour)
    miss = []
    for name, e in surface[flavour].items():
        sigs = ours.get(name)
        if not sigs:
            miss.append((name, e["arities"], "not declared")); continue
        for a in e["arities"]:
            if not any(lo <= a <= hi for lo, hi in sigs):
                miss.append((name, [a], "no overload taking %d arguments (have %s)" % (a, sigs)))
    return miss


if __name__ == "__main__":
    surface = scan_reference()
    out = os.path.join(REPO, "tests", "golden", "dropin_surface.json")
    json.dump(surface, open(out, "w"), indent=1, sort_keys=True)
    for flav in surface:
        print(flav, len(surface[flav]), "members called by the planners:", " ".join(sorted(surface[flav])))
        for m in missing(surface, flav):
            print("   MISSING", m)
