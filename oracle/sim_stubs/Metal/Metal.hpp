// Stand-in for Apple's metal-cpp umbrella header: the reference's include/systems/fluid/fluid.hpp only names these
// types as pointer members (fluid.hpp:328-349); the headless harness never constructs them.
#pragma once
namespace MTL {
class Device;
class CommandQueue;
class Library;
class ComputePipelineState;
class ComputeCommandEncoder;
class Buffer;
}  // namespace MTL
