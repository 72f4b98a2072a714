// TEST INFRASTRUCTURE ONLY (oracle/): declaration-level stand-in for <tbb/enumerable_thread_specific.h>; src/ml_optimiser.h
// only holds one as a member (constructed from an initial value) and never uses it in the files built here.
#pragma once
namespace tbb {
template <typename T> class enumerable_thread_specific {
	T init_;
public:
	enumerable_thread_specific() : init_() {}
	explicit enumerable_thread_specific(T v) : init_(v) {}
	T &local() { static thread_local T t = init_; return t; }
};
}
