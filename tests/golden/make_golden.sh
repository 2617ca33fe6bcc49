#!/bin/sh
# How tests/golden/data was made (run in the build container, where /root/reference is mounted):
# the reference's generator, unmodified, in a scratch directory; it writes ../data relative to its cwd.
set -e
rm -rf /tmp/caf_golden && mkdir -p /tmp/caf_golden/utils
cp /root/reference/utils/generate.py /tmp/caf_golden/utils/
(cd /tmp/caf_golden/utils && python generate.py)
mkdir -p "$(dirname "$0")/data"
cp /tmp/caf_golden/data/*.c64 "$(dirname "$0")/data/"
# known_answers.json restates the assert_eq! lines of /root/reference/caf_rust/tests/test.rs by hand.
