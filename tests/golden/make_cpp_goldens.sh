#!/bin/bash
# Golden hashes for the stand-alone value types, made from the UNMODIFIED reference: the test
# programs under tests/cpp built against the reference's own headers and sources (needs
# /root/reference; run from the repo root after `make -C perturb_b200/csrc`).
set -e
REF=${REF:-/root/reference}
SRC="$REF/src/perturb.cpp $REF/src/sgp4.cpp $REF/src/tle.cpp"
g++ -std=c++11 -w -DPERTURB_BATCH_USE_PERTURB_HPP -I$REF/include -Iinclude tests/cpp/julian_dates.cpp $SRC -o /tmp/julian_dates_ref
/tmp/julian_dates_ref | sha256sum | cut -d' ' -f1 > tests/golden/julian_dates.sha256
/tmp/julian_dates_ref | wc -l >> tests/golden/julian_dates.sha256
g++ -std=c++11 -w -DPERTURB_BATCH_USE_PERTURB_HPP -I$REF/include -Iinclude tests/cpp/tle_members.cpp $SRC -o /tmp/tle_members_ref
python -c "import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, '.'); import helpers; sys.stdout.buffer.write(helpers.tle_parser_feed())" > /tmp/tle_feed.txt
/tmp/tle_members_ref < /tmp/tle_feed.txt | sha256sum | cut -d' ' -f1 > tests/golden/tle_members.sha256
/tmp/tle_members_ref < /tmp/tle_feed.txt | wc -l >> tests/golden/tle_members.sha256
/tmp/tle_members_ref < /tmp/tle_feed.txt | grep "^err" | sort | uniq -c | tr -s ' ' | tr '\n' ';' >> tests/golden/tle_members.sha256
echo >> tests/golden/tle_members.sha256
cat tests/golden/julian_dates.sha256 tests/golden/tle_members.sha256
