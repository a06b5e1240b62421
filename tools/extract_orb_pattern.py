"""Extract the 256x4 learned rBRIEF sampling pattern (a DATA table that OpenCV
publishes as bit_pattern_31_; the reference carries a copy at
src/ORBextractor.cc:150-408) into a bare comma-separated include file.

Run once in the build container (the reference tree is absent on GPU boxes);
the output is committed.
"""
import re, sys, pathlib
src = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src/ORBextractor.cc").read_text()
m = re.search(r"bit_pattern_31_\[256\*4\]\s*=\s*\{(.*?)\};", src, re.S)
body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
nums = [int(t) for t in re.findall(r"-?\d+", body)]
assert len(nums) == 1024, len(nums)
out = pathlib.Path(__file__).resolve().parent.parent / "pl-slam_b200" / "data" / "orb_pattern_31.inc"
lines = ["// 256 test pairs (x0,y0,x1,y1) of the steered-BRIEF pattern, patch 31; data only.\n"]
for i in range(0, 1024, 16):
    lines.append(",".join(str(v) for v in nums[i:i+16]) + ",\n")
out.write_text("".join(lines))
print("wrote", out, max(nums), min(nums))
