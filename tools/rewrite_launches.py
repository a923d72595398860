"""One-off source transformation: kernel<<<grid, block, smem, stream>>>(args) -> caae::launch(kernel, grid, block, smem,
stream, args) (programmatic dependent launch, common.cuh) and `pdl_wait();` as the first statement of every __global__
kernel.  Idempotent."""
import re, sys

def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{<" and not (ch == "<" and False): depth += ch in "([{"
        if ch in ")]}": depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out

def rewrite(src):
    i = 0
    res = ""
    n = 0
    while True:
        k = src.find("<<<", i)
        if k < 0:
            res += src[i:]; break
        # kernel expression: walk back over identifier chars / template args
        j = k
        depth = 0
        while j > 0:
            c = src[j - 1]
            if c == ">": depth += 1
            elif c == "<": depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"): break
            j -= 1
        kern = src[j:k]
        e = src.find(">>>", k)
        cfg = split_top(src[k + 3:e])
        # argument list
        assert src[e + 3] == "(", (kern, src[e:e + 20])
        d, m = 0, e + 3
        while True:
            if src[m] == "(": d += 1
            elif src[m] == ")":
                d -= 1
                if d == 0: break
            m += 1
        args = src[e + 4:m]
        while len(cfg) < 4: cfg.append("0")
        call = f"caae::launch({kern}, {cfg[0]}, {cfg[1]}, {cfg[2]}, {cfg[3]}" + (", " + args if args.strip() else "") + ")"
        res += src[i:j] + call
        i = m + 1
        n += 1
    return res, n

def add_waits(src):
    out, pos, n = "", 0, 0
    for mt in re.finditer(r"__global__", src):
        b = src.find("{", mt.end())
        semi = src.find(";", mt.end())
        if semi != -1 and semi < b: continue          # a declaration
        if src[b + 1:b + 40].lstrip().startswith("pdl_wait();"): continue
        out += src[pos:b + 1] + "\n  pdl_wait();"
        pos = b + 1
        n += 1
    return out + src[pos:], n

for path in sys.argv[1:]:
    s = open(path).read()
    s, a = rewrite(s)
    s, b = add_waits(s)
    open(path, "w").write(s)
    print(path, a, "launches,", b, "kernels")
