#!/usr/bin/env python
"""Write whisper_b200/assets/checkpoints.json from the reference's own tables (run in the build container, where the
reference is importable from /root/reference): for every official model name the file name and SHA-256 digest the
reference expects in its cache directory (whisper/__init__.py:17-32: both are the last two components of the download
URL) and the base85 dump of its word-timing alignment heads (whisper/__init__.py:36-51).  Data tables only."""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
import whisper  # noqa: E402  (the reference)

out = {}
for name, url in whisper._MODELS.items():
    out[name] = {"file": os.path.basename(url), "sha256": url.split("/")[-2],
                 "alignment_heads": whisper._ALIGNMENT_HEADS[name].decode("ascii") if name in whisper._ALIGNMENT_HEADS else None}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "whisper_b200", "assets", "checkpoints.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print(f"wrote {path}: {len(out)} models")
