#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3: nothing charged).  usage: tools/gpurun_retry.sh <gpurun args...>
for i in $(seq 1 20); do
	/usr/local/graft/bin/gpurun "$@"
	rc=$?
	if [ $rc -ne 3 ]; then exit $rc; fi
	sleep 90
done
exit 3
