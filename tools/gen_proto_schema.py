"""Generates mtl_ssl_b200/protos/schema.json from the reference's proto2 schema files.

Run in the build container (where /root/reference exists):
    python tools/gen_proto_schema.py /root/reference/object_detection/protos
The GPU box has no /root/reference, so the derived JSON (message -> fields with label, type,
default; enums -> value names) is committed and loaded by mtl_ssl_b200/protos/text_format.py.
Only the schema FACTS travel (field names, numbers, defaults); no reference source is copied.
"""
import json
import os
import re
import sys


def strip_comments(s):
    return re.sub(r"//[^\n]*", "", s)


def parse_block(body, prefix, messages, enums):
    """body: text inside a message {...}; registers nested messages/enums, returns field list."""
    fields = []
    i = 0
    n = len(body)
    oneof = None
    while i < n:
        m = re.compile(r"\s*(message|enum|oneof)\s+(\w+)\s*\{").match(body, i)
        if m:
            kind, name = m.group(1), m.group(2)
            depth, j = 1, m.end()
            while depth:
                if body[j] == "{":
                    depth += 1
                elif body[j] == "}":
                    depth -= 1
                j += 1
            inner = body[m.end():j - 1]
            if kind == "message":
                messages[prefix + name] = parse_block(inner, prefix + name + ".", messages, enums)
            elif kind == "enum":
                enums[prefix + name] = {k: int(v) for k, v in re.findall(r"(\w+)\s*=\s*(-?\d+)\s*;", inner)}
            else:
                for f in parse_block(inner, prefix, messages, enums):
                    f["oneof"] = name
                    fields.append(f)
            i = j
            continue
        m = re.compile(r"\s*(optional|repeated|required)?\s*([\w.]+)\s+(\w+)\s*=\s*(\d+)\s*(\[[^\]]*\])?\s*;").match(body, i)
        if m:
            label, typ, name, num, opts = m.groups()
            f = {"name": name, "type": typ, "label": label or "optional", "number": int(num)}
            if opts:
                d = re.search(r"default\s*=\s*([^,\]]+)", opts)
                if d:
                    f["default"] = d.group(1).strip().strip("'\"")
            fields.append(f)
            i = m.end()
            continue
        m = re.compile(r"\s*[^;{}]*;").match(body, i)      # syntax/package/import/option lines
        if m and m.end() > i:
            i = m.end()
            continue
        i += 1
    return fields


def main(proto_dir, out_path):
    messages, enums = {}, {}
    for fn in sorted(os.listdir(proto_dir)):
        if fn.endswith(".proto"):
            parse_block(strip_comments(open(os.path.join(proto_dir, fn)).read()), "", messages, enums)
    json.dump({"messages": messages, "enums": enums}, open(out_path, "w"), indent=0, sort_keys=True)
    print("wrote %s: %d messages, %d enums" % (out_path, len(messages), len(enums)))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/object_detection/protos",
         os.path.join(here, "..", "mtl_ssl_b200", "protos", "schema.json"))
