#!/bin/bash
# usage: tools/ab_snapshot.sh [rev=HEAD]  -> build/ab_base2 = a built copy of the repo at <rev> (for tools/ab_dirs.sh)
REV=${1:-HEAD}
ROOT=$(cd $(dirname $0)/.. && pwd)
W=/tmp/bt_ab_worktree
if [ ! -d $W ]; then git -C $ROOT worktree add -f --detach $W $REV >/dev/null 2>&1; fi
( cd $W && git checkout -q --detach $(git -C $ROOT rev-parse $REV) && python -m brax_tracking_b200.build >/dev/null ) || exit 1
rm -rf $ROOT/build/ab_base2 && mkdir -p $ROOT/build/ab_base2
( cd $W && tar cf - --exclude=.git --exclude=gpurun_out --exclude=profiles --exclude='brax_tracking_b200/build' . ) | ( cd $ROOT/build/ab_base2 && tar xf - )
ls -la $ROOT/build/ab_base2/brax_tracking_b200/libbt_b200.so
