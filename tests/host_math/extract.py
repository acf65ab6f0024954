"""Pull named functions / constants out of the CUDA headers as TEXT so a host C++ program can compile the very same
source (tests/test_kernel_math_host.py).  The headers themselves are not touched: the device build is unchanged."""
import re


def _balanced(src: str, start: int) -> int:
    """index one past the '}' (or ';' for declarations without a body) that closes the item starting at `start`"""
    i = src.index("{", start)
    semi = src.find(";", start)
    if semi != -1 and semi < i:
        return semi + 1
    depth = 0
    while True:
        c = src[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                i += 1
                # struct / array initialiser ends with ';'
                j = i
                while j < len(src) and src[j] in " \t":
                    j += 1
                return j + 1 if j < len(src) and src[j] == ";" else i
        i += 1


def extract(src: str, names) -> str:
    """every definition whose declarator mentions `name(` (functions, all overloads / templates) or `name` followed by
    '=' / '[' / '{' (constants, arrays, structs), in source order, with its leading `template <...>` line if present"""
    spans = []
    for name in names:
        pat = re.compile(r"^(?:template\s*<[^>]*>\s*\n)?[^\n;{}()]*\b" + re.escape(name) + r"\b\s*(?:\(|=|\[|\{)", re.M)
        found = False
        for m in pat.finditer(src):
            line = src[m.start():src.index("\n", m.start())]
            if line.lstrip().startswith("//"):
                continue
            spans.append((m.start(), _balanced(src, m.start()) if "{" in src[m.start():src.find(";", m.start()) + 1] or
                          "(" in m.group(0) or "struct" in m.group(0) else src.index(";", m.start()) + 1))
            found = True
        if not found:
            raise KeyError(f"{name} not found in the header")
    spans = sorted(set(spans))
    out, last = [], -1
    for a, b in spans:
        if a >= last:
            out.append(src[a:b])
            last = b
    return "\n".join(out) + "\n"
