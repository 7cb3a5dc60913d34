"""Documentation integrity: every `profiles/...` file and every repo path that DESIGN.md, INTEGRATION.md, README.md or
profiles/README.md name exists, and every C entry point the header declares is mentioned by name in the header's own
comment or the docs' tables is at least exported (tests/test_abi.py) - stale references are the first thing a reader trips
over."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _text(name):
    return open(os.path.join(ROOT, name)).read()


def test_referenced_profile_files_exist():
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md")):
        txt = _text(doc)
        names = set(re.findall(r"`((?:profiles/)?r01_[a-z0-9]+_[A-Za-z0-9_{},.*]+?\.(?:md|json|csv|txt|log))`", txt))
        for n in names:
            base = n.split("/")[-1]
            if "*" in base or "{" in base:
                pat = re.escape(base).replace(r"\*", ".*")
                pat = re.sub(r"\\\{([^}]*)\\\}", lambda m: "(" + "|".join(m.group(1).replace("\\", "").split(",")) + ")", pat)
                if not any(re.fullmatch(pat, f) for f in os.listdir(os.path.join(ROOT, "profiles"))):
                    missing.append((doc, n))
            elif not os.path.exists(os.path.join(ROOT, "profiles", base)):
                missing.append((doc, n))
    assert not missing, missing


def test_referenced_source_paths_exist():
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md"):
        for n in set(re.findall(r"`((?:copo_b200|tests|tools|oracle|include)/[A-Za-z0-9_/.]+\.(?:py|cu|cuh|h|sh|npz|cpp))`", _text(doc))):
            if not os.path.exists(os.path.join(ROOT, n)):
                missing.append((doc, n))
    assert not missing, missing
