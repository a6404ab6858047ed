/* TEST INFRASTRUCTURE ONLY (oracle/): force-included before the reference's
 * unmodified statePropagator.cu / collisionCheck.cu so g++ can build them for
 * the host.  Erases the CUDA execution-space qualifiers; nothing else. */
#pragma once
#define __device__
#define __host__
#define __global__
