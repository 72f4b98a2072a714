// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for <tbb/spin_mutex.h>.
// TBB is not installed in this image; the reference's ALTCPU back-projection kernels
// (src/acc/cpu/cpu_kernels/BP.h) only use tbb::spin_mutex + scoped_lock.
#pragma once
#include <atomic>

namespace tbb {
class spin_mutex {
	std::atomic_flag f = ATOMIC_FLAG_INIT;
public:
	void lock()   { while (f.test_and_set(std::memory_order_acquire)) { } }
	void unlock() { f.clear(std::memory_order_release); }
	class scoped_lock {
		spin_mutex &m;
	public:
		explicit scoped_lock(spin_mutex &mm) : m(mm) { m.lock(); }
		~scoped_lock() { m.unlock(); }
	};
};
}
